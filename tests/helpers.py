"""Shared test helpers: build the float64 oracle and the CUDA inputs from one scene."""

import numpy as np
import torch

from oracle import boundaries as ob
from oracle import engine as oe
from oracle import materials as om
from tensorflowraytrace_b200 import scenes

TRI_FIELDS = ("xp", "yp", "zp", "x1", "y1", "z1", "x2", "y2", "z2")
RAY_FIELDS = ("x_start", "y_start", "z_start", "x_end", "y_end", "z_end")

# fp32-appropriate epsilons shared by oracle and kernel in parity tests
EPS = dict(intersect_epsilion=1e-10, size_epsilion=1e-5, ray_start_epsilion=1e-4)


def f32_round(x):
    """Round float64 data through float32: the oracle then sees exactly the kernel's inputs."""
    return np.asarray(x, dtype=np.float32).astype(np.float64)


def oracle_from_scene(scene, ray_slice=None, requires_grad=False, eps=EPS, round_inputs=True,
                      **engine_kwargs):
    """Returns (engine, parameters list).  Inputs are rounded through fp32 (vertices are formed in
    fp64 from fp32-rounded zero points / vectors / parameters, then the *fields* are what they
    are: the kernel computes vertices in fp32, see `round_vertices`)."""
    rnd = f32_round if round_inputs else (lambda a: np.asarray(a, dtype=np.float64))
    src = {}
    for k, v in scene["sources"].items():
        v = v if ray_slice is None else v[ray_slice]
        src[k] = torch.from_numpy(np.ascontiguousarray(rnd(v)))
    params, optical = [], []
    for s in scene["surfaces"]:
        p = torch.from_numpy(rnd(s["parameters"])).clone().requires_grad_(requires_grad)
        params.append(p)
        verts = ob.parametric_vertices(torch.from_numpy(rnd(s["zero_points"])),
                                       torch.from_numpy(rnd(s["vectors"])), p)
        if round_inputs and not requires_grad:
            verts = torch.from_numpy(f32_round(verts.detach().numpy()))
        fields = ob.fields_from_vertices(verts, s["faces"], s.get("vertex_update_map"))
        F = s["faces"].shape[0]
        fields["mat_in"] = torch.full((F,), s["mat_in"], dtype=torch.int64)
        fields["mat_out"] = torch.full((F,), s["mat_out"], dtype=torch.int64)
        optical.append(fields)

    def static(lst):
        out = []
        for m in lst:
            out.append(ob.fields_from_vertices(torch.from_numpy(rnd(m["vertices"])), m["faces"]))
        return out

    system = oe.OracleSystem3D(
        [src], optical, static(scene["stops"]), static(scene["targets"]),
        materials=[{"n": om.BY_NAME[name]} for name in scene["materials"]], **eps)
    inherit = tuple(k for k in scene["sources"] if k not in RAY_FIELDS)
    engine = oe.OracleEngine3D(system, simple_ray_inheritance=inherit, **engine_kwargs)
    return engine, params


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = max(np.max(np.abs(b)) if b.size else 0.0, 1e-30)
    return float(np.max(np.abs(a - b)) / scale) if a.size else 0.0
