#!/bin/bash
# AddressSanitizer run of the single-lane host build of the kernel source (tests/hostsim) over every fixture shape
# (trajectory and path mode): the host buffers are sized by the same tg_sqp_workspace_doubles / tg_scratch_doubles
# the CUDA library uses, so an out-of-bounds index of tg_eval.h / tg_sqp.h shows up here.  CPU only.
# Two builds: the fused variant of the QP stage and the lock-step kernels' one (-DTG_FUSED_LM_FAR: the factor is updated
# on a copy that spans J's storage, R's and the vectors behind them).
set -e
cd "$(dirname "$0")/.."
for VARIANT in "" "-DTG_FUSED_LM_FAR"; do
echo "== host build ${VARIANT:-(fused variant)}"
g++ -std=c++14 -O1 -g -fPIC -shared -fsanitize=address -fno-omit-frame-pointer -ffp-contract=off -Wno-unknown-pragmas \
    -DTG_WITH_SQP $VARIANT -o /tmp/asan_hostsim.so tests/hostsim/tg_hostsim.cpp
LD_PRELOAD=$(g++ -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python - <<'PY'
import sys
sys.path[:0] = [".", "tests", "oracle"]
import helpers, path_problems, problems, hostsim_loader
from trajectory_generator_b200.problem import pack_problem
hs = hostsim_loader.HostSim("/tmp/asan_hostsim.so")
ns = helpers.product_namespace()
for table, default, mode in ((path_problems.ALL, "minimal_velocity_path", True), (problems.ALL, "minimal_velocity_and_time_path", False)):
    for name, make in table.items():
        d, cc, kw = make(ns)
        pm = ("indirect" if kw.get("isIndirect") else "direct") if mode else None
        pp = pack_problem(d, cc, kw.get("objective_function_type", default), kw.get("num_intervals_free_space"), path_mode=pm)
        for fd in (True, False):
            r = hs.solve(pp, fd=fd, maxiter=40)
        hs.eval(pp, pp.x0)
        hs.fd_derivatives(pp, pp.x0)
        print("%-5s %-28s clean (status %d, %d iterations)" % ("path" if mode else "traj", name, r["status"], r["nit"]))
# the spline order converter (csrc/tg_smooth.h)
import ctypes, numpy as np
import tg_oracle_smoothing as osm
ND = np.ctypeslib.ndpointer(dtype=np.float64, flags="C")
hs.lib.hs_smooth_solve.argtypes = [ctypes.c_int] * 4 + [ctypes.c_double, ND, ND, ND, np.ctypeslib.ndpointer(dtype=np.int32, flags="C")]
hs.lib.hs_smooth_initial.argtypes = [ctypes.c_int, ND, ctypes.c_int, ctypes.c_int, ND]
for name, g in helpers.load_golden("smoothing.json")["cases"].items():
    cp = np.array(g["cp"], dtype=float)
    prob = osm.SmoothingProblem(g["new_order"], cp, g["scale"], g["old_order"], g["resolution"])
    x0 = np.zeros((prob.d, prob.N)); hs.lib.hs_smooth_initial(prob.d, np.ascontiguousarray(cp), cp.shape[1], prob.N, x0)
    x = x0.flatten().copy(); f = np.zeros(1); nit = np.zeros(1, np.int32)
    st = hs.lib.hs_smooth_solve(prob.d, prob.N, prob.k, g["resolution"], prob.scale, np.concatenate([prob.Y.flatten(), prob.b.flatten()]), x, f, nit)
    print("smooth %-26s clean (status %d, %d iterations)" % (name, st, nit[0]))
PY
done
