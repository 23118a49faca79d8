"""Imaging-lens optimisation on the GPU, after the reference's dev/simple_3d_optimize.py:22-290.

A two-surface acrylic lens (ParametricMultiTriangleBoundary over a hexagonal disk, thickness
constraints, vertex_update_map + gradient accumulator from mesh_tools) is optimised with
SGD_Optimizer so that random rays from a 0.4 x 0.4 object at x = -4 form a 2x magnified, inverted
image at x = 8.  Every step regenerates the random source rays on the device, traces them with
the fused CUDA kernels, back-propagates to the surface parameters and applies the update.

    python examples/simple_3d_optimize.py [--rays 20000] [--rings 12] [--steps 60]
"""
import argparse
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorflowraytrace_b200 import (boundaries, distributions, engine, materials,  # noqa: E402
                                     mesh_tools, operation, optimizer, scenes, sources)

SOURCE_DISTANCE, MAGNIFICATION, OBJECT_SIZE = 4.0, 2.0, 0.2


def build(ray_count=20000, rings=12, seed=0, device="cuda", grad_clip=1e-3):
    # random source (dev/simple_3d_optimize.py:29-43): re-sampled on the device at every update
    distributions.seed_all(seed)
    side = max(int(round(ray_count ** 0.5)), 1)
    base_points = distributions.RandomUniformSquare(OBJECT_SIZE, side)
    angles = distributions.RandomUniformSphere(math.pi / 16.0, side * side)
    source = sources.AngularSource(
        3, (-SOURCE_DISTANCE, 0.0, 0.0), (1.0, 0.0, 0.0), angles, base_points, [scenes.YELLOW],
        dense=False, extra_fields={"rank": ("base_point", base_points, "ranks")})

    # zero points: a hexagonal disk in the y-z plane
    pts, faces = mesh_tools.hexagonal_mesh(1.0, rings)
    zero = boundaries.TriangleMesh(np.stack([np.zeros(len(pts)), pts[:, 0], pts[:, 1]], 1), faces)
    top_parent = mesh_tools.get_closest_point(zero, (0, 0, 0))
    vertex_update_map, accumulator = mesh_tools.mesh_parametrization_tools(zero, top_parent)
    smoother = mesh_tools.mesh_smoothing_tool(zero, mesh_tools.gaussian_weights(1.0, 3))

    lens = boundaries.ParametricMultiTriangleBoundary(
        zero, boundaries.FromVectorVG((1, 0, 0)),
        [boundaries.ThicknessConstraint(0.0, "min"), boundaries.ThicknessConstraint(0.2, "min")],
        [True, False], material_list=[{"mat_in": np.int64(1), "mat_out": np.int64(0)}] * 2,
        vertex_update_map=vertex_update_map)
    tv, tf_ = mesh_tools.plane_mesh((SOURCE_DISTANCE * MAGNIFICATION, 0.0, 0.0), 50.0, axis=0)
    target = boundaries.ManualTriangleBoundary(mesh=boundaries.TriangleMesh(tv, tf_))
    target.frozen = True

    system = engine.OpticalSystem3D()
    system.optical = lens.surfaces
    system.targets = [target]
    system.sources = [source]
    system.materials = [{"n": materials.vacuum}, {"n": materials.acrylic}]
    system.update()

    trace_engine = engine.OpticalEngine(3, [operation.StandardReaction()],
                                        simple_ray_inheritance={"wavelength", "rank"},
                                        compile_active_rays=False)
    trace_engine.optical_system = system
    trace_engine.validate_system()

    def error_function(eng):
        fin = eng.finished_rays
        goal = -(MAGNIFICATION * OBJECT_SIZE) * fin["rank"]
        return (fin["y_end"] - goal[:, 0]) ** 2 + (fin["z_end"] - goal[:, 1]) ** 2

    dev = torch.device(device)
    acc = mesh_tools.to_torch_sparse(accumulator, dev)
    smo = mesh_tools.to_torch_sparse(smoother, dev)
    opt = optimizer.SGD_Optimizer(trace_engine, lens.parameters, error_function, 3,
                                  learning_rate=1.0, grad_clip=grad_clip, sgd_lr=1.0)
    return trace_engine, lens, opt, acc, smo


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=20000)
    ap.add_argument("--rings", type=int, default=12)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--lr", type=float, default=2e-5)
    ap.add_argument("--clip", type=float, default=1e-3)
    ap.add_argument("--momentum", type=float, default=0.8)
    args = ap.parse_args()
    trace_engine, lens, opt, acc, smo = build(args.rays, args.rings, grad_clip=args.clip)
    routine = [
        {"steps": args.steps // 2, "learning_rate": (args.lr, args.lr / 2), "momentum": args.momentum,
         "accumulators": acc, "smoothers": smo},
        {"steps": args.steps - args.steps // 2, "learning_rate": args.lr / 4,
         "momentum": args.momentum / 2},
    ]
    opt.training_routine(routine, report_frequency=max(args.steps // 10, 1))
    trace_engine.optical_system.update()     # re-applies the thickness constraints
    thick = (lens.parameters[1] - lens.parameters[0]).detach()
    print("thickness range:", float(thick.min()), float(thick.max()))


if __name__ == "__main__":
    main()
