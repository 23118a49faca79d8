"""
SGD_Optimizer: the caller of the fwd+bwd hot path (tfrt/optimizer.py:14-442).

`process_gradient` is the reference's gradient seam (tfrt/optimizer.py:187-258): update the
system, trace, evaluate the user's error function on the engine, back-propagate to the boundary
parameters, then post-process (finite filter, learning rate, clip, accumulator).  Here the tape
is torch autograd over two CUDA nodes (autograd.FacesFunction, autograd.TraceFunction); when the
process group is initialised with more than one rank the parameter gradients and the error sum
are all-reduced (parallel.py) before post-processing, so every rank applies the same step.

The parameter update is Keras' SGD with Nesterov momentum (the reference instantiates
`tf.optimizers.SGD(nesterov=True)` and scales gradients itself, tfrt/optimizer.py:103, :243-250,
:316), restated: v <- m v - lr g ; p <- p + m v - lr g (Nesterov) with lr = `sgd_lr`.
"""

import time

import torch

from . import parallel


class SGD_Optimizer:
    def __init__(self, engine, parameters, error_function, trace_depth, momentum=0.0,
                 learning_rate=1.0, individual_lr=None, grad_clip="default", clip_mode="common",
                 clip_scale=10.0, sgd_lr=0.01):
        self.engine = engine
        if type(parameters) is list or type(parameters) is tuple:
            self.parameters = parameters
        else:
            raise ValueError("SGD_Optimizer: parameters must be a list of tf.variable")
        self.error_function = error_function
        self.trace_depth = trace_depth
        self.sgd_lr = sgd_lr            # tf.optimizers.SGD default learning_rate
        self._velocity = [None] * len(parameters)
        self.momentum = momentum
        self.learning_rate = learning_rate
        self.individual_lr = individual_lr
        self.clip_scale = clip_scale
        self.grad_clip = self.clip_scale * learning_rate if grad_clip == "default" else grad_clip
        self.clip_mode = clip_mode
        self.suppress_warnings = False
        self.iterations = 0

    @property
    def momentum(self):
        return self._momentum

    @momentum.setter
    def momentum(self, val):
        if 0.0 <= val <= 1.0:
            self._momentum = val
        else:
            raise ValueError("SGD_Optimizer: Momentum must be between 0 and 1.")

    @property
    def individual_lr(self):
        return self._individual_lr

    @individual_lr.setter
    def individual_lr(self, val):
        if val is None:
            self._individual_lr = [1.0] * len(self.parameters)
            return
        try:
            if len(val) != len(self.parameters):
                raise ValueError(
                    "SGD_Optimizer: individual_lr must have as many elements as there are "
                    "parameters.")
        except TypeError as e:
            raise TypeError(
                "SGD_Optimizer: individual_lr must have as many elements as there are "
                "parameters.") from e
        self._individual_lr = val

    def convert_to_plist(self, data):
        p_count = len(self.parameters)
        if type(data) is list or type(data) is tuple:
            if len(data) == p_count:
                return data
            raise ValueError("SGD_Optimizer: plist arguments must have one element per parameter.")
        return [data] * p_count

    def convert_to_lrlist(self, lr, steps):
        try:
            return torch.linspace(float(lr[0]), float(lr[1]), steps).tolist()
        except TypeError:
            return [lr] * steps

    # ------------------------------------------------------------------ the gradient seam
    def raw_gradient(self, *args, **kwargs):
        """system.update() -> ray_trace -> error -> d(sum error)/d(parameters).  Returns
        (grads list with None for unreachable parameters, error tensor).  With several ranks
        the gradients and the error are NOT yet reduced (see process_gradient)."""
        self.engine.clear_ray_history()
        self.engine.optical_system.update()
        self.engine.ray_trace(self.trace_depth)
        error = self.error_function(self.engine, *args, **kwargs)
        total = error.sum()
        if total.requires_grad:
            grads = torch.autograd.grad(total, self.parameters, allow_unused=True)
        else:
            grads = [None] * len(self.parameters)
        return list(grads), error

    def process_gradient(self, accumulators, *args, lr_scale=1.0, **kwargs):
        """Compute and process the gradients (tfrt/optimizer.py:187-258).  Returns
        (processed grads, mean error)."""
        grads, error = self.raw_gradient(*args, **kwargs)
        err_sum, err_count = error.sum().detach(), float(error.numel())
        if parallel.is_distributed():
            # rays are sharded: a rank that reached no parameter still takes part in the reduce
            grads = [torch.zeros_like(p) if g is None else g.contiguous()
                     for g, p in zip(grads, self.parameters)]
            stats = torch.stack([err_sum.to(torch.float32),
                                 torch.tensor(err_count, device=err_sum.device)])
            parallel.allreduce_sum_(grads + [stats])
            err_sum, err_count = stats[0], float(stats[1].item())
        mean_error = err_sum / max(err_count, 1.0)

        processed = []
        for i, params in enumerate(self.parameters):
            grad = grads[i]
            if grad is None:
                if not self.suppress_warnings:
                    print(
                        "Warning: SGD_Optimizer.process_gradient encountered a possible issue:  "
                        "The gradient was None, which can mean that the error does not depend "
                        "on it.  The gradient will be set to zero and future instances of this "
                        "message will be suppressed.")
                    self.suppress_warnings = True
                grad = torch.zeros_like(params)
            else:
                grad = torch.where(torch.isfinite(grad), grad, torch.zeros_like(grad))
            grad = grad * (lr_scale * self.individual_lr[i] * self.learning_rate)
            if self.clip_mode == "common":
                grad = torch.clamp(grad, -self.grad_clip, self.grad_clip)
            else:
                clp = self.individual_lr[i] * self.clip_scale * self.learning_rate * lr_scale
                grad = torch.clamp(grad, -clp, clp)
            if accumulators[i] is not None:
                acc = accumulators[i]
                acc = acc if torch.is_tensor(acc) else torch.as_tensor(acc)
                acc = acc.to(device=grad.device, dtype=grad.dtype)
                prev_shape = grad.shape
                if acc.is_sparse:
                    grad = torch.sparse.mm(acc, grad.reshape(-1, 1)).reshape(prev_shape)
                else:
                    grad = (acc @ grad.reshape(-1, 1)).reshape(prev_shape)
            processed.append(grad)
        return processed, mean_error

    @staticmethod
    def smooth(parameters, smoother):
        """parameters <- smoother @ parameters (tfrt/optimizer.py:260-282)."""
        if smoother is not None:
            s = smoother if torch.is_tensor(smoother) else torch.as_tensor(smoother)
            s = s.to(device=parameters.device, dtype=parameters.dtype)
            with torch.no_grad():
                flat = parameters.reshape(-1, 1)
                new = torch.sparse.mm(s, flat) if s.is_sparse else s @ flat
                parameters.copy_(new.reshape(parameters.shape))

    def apply_gradients(self, grads):
        """Keras SGD(nesterov=True) update with the current momentum."""
        m, lr = self._momentum, self.sgd_lr
        with torch.no_grad():
            for i, (g, p) in enumerate(zip(grads, self.parameters)):
                if m > 0.0:
                    v = self._velocity[i]
                    if v is None:
                        v = torch.zeros_like(p)
                    v = m * v - lr * g
                    self._velocity[i] = v
                    p.add_(m * v - lr * g)
                else:
                    p.add_(-lr * g)

    def single_step(self, accumulators, *args, lr_scale=1.0, momentum=0.0, verbose=False,
                    **kwargs):
        """One optimisation step (tfrt/optimizer.py:284-320); returns the mean error."""
        self.momentum = momentum
        grads, error = self.process_gradient(accumulators, *args, lr_scale=lr_scale, **kwargs)
        self.apply_gradients(grads)
        self.iterations += 1
        error = float(error)
        if verbose:
            print(f"step {self.iterations} error: {error}")
        return error

    def training_routine(self, routine, post_step=None, report_frequency=1, show_time=True):
        """Run the phases of `routine` (tfrt/optimizer.py:322-442).  Each phase dict updates the
        running hyper-parameters {steps, learning_rate (relative; float or (start, end)),
        momentum, accumulators, smoothers, erf_args, erf_kwargs, individual_lr}; keys a phase
        does not name keep the previous phase's value."""
        phase = {"steps": 10, "learning_rate": 1.0, "momentum": 0.0, "accumulators": None,
                 "smoothers": None, "erf_args": [], "erf_kwargs": {}, "individual_lr": None}
        self.iterations = 0
        start_time = time.time()
        steps, total_iterations = phase["steps"], 0
        for new_phase in routine:
            steps = new_phase.get("steps", steps)
            total_iterations += steps
        for current_phase, new_phase in enumerate(routine, start=1):
            phase.update(new_phase)
            phase["accumulators"] = self.convert_to_plist(phase["accumulators"])
            phase["smoothers"] = self.convert_to_plist(phase["smoothers"])
            lrs = self.convert_to_lrlist(phase["learning_rate"], phase["steps"])
            self.individual_lr = phase["individual_lr"]
            for i in range(phase["steps"]):
                error = self.single_step(
                    phase["accumulators"], *phase["erf_args"], lr_scale=lrs[i],
                    momentum=phase["momentum"], verbose=False, **phase["erf_kwargs"])
                for p, s in zip(self.parameters, phase["smoothers"]):
                    self.smooth(p, s)
                if report_frequency != 0 and self.iterations % report_frequency == 0:
                    print(f"Phase {current_phase}/{len(routine)}, step {i + 1}/{phase['steps']}, "
                          f"total {self.iterations}/{total_iterations}-"
                          f"{100 * self.iterations / total_iterations:.1f}%.  Error: {error}.")
                if post_step:
                    post_step()
        total_time = time.time() - start_time
        if show_time:
            print(f"Completed training routine.  Took {total_time} seconds.")
            print(f"Steps took an average of {total_time / max(total_iterations, 1)} seconds "
                  "per step.")
