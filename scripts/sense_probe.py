"""Device-timed measurement of the local-map acquisition kernel without torch (ctypes on libcudart + the C ABI):
python scripts/sense_probe.py [n_agents] [steps] [cpu_agents] -> one JSON object on stdout.
Same workload as bench.py's sense_measure (one shared forest environment grid, 66x66x20 local grids, steady-state
update with kept grids; L2 flushed by a 512 MiB memset between timed launches).  HDSM_SENSE_BITS=0 in the environment
selects the kernel's first (key) form for an A/B against the default bitmap form."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_agent_pkgs_b200 import scenarios as sc, sensing as sn  # noqa: E402

rt = C.CDLL("libcudart.so.12")


def ck(rc):
    if rc != 0:
        raise RuntimeError(f"CUDA runtime error {rc}")


def dmalloc(nbytes):
    p = C.c_void_p()
    ck(rt.cudaMalloc(C.byref(p), C.c_size_t(nbytes)))
    return p


def h2d(dst, arr):
    arr = np.ascontiguousarray(arr)
    ck(rt.cudaMemcpy(dst, arr.ctypes.data_as(C.c_void_p), C.c_size_t(arr.nbytes), C.c_int(1)))


def d2h(arr, src):
    ck(rt.cudaMemcpy(arr.ctypes.data_as(C.c_void_p), src, C.c_size_t(arr.nbytes), C.c_int(2)))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    m = min(n, int(sys.argv[3]) if len(sys.argv) > 3 else 512)
    rng = np.random.default_rng(4)
    vox, rng3 = 0.3, (20.0, 20.0, 6.0)
    world = sc.Forest.density(rng, (0.0, 0.0), (114.0, 114.0), 0.2)
    env, org = sn.environment_grid(world, vox)
    pos0 = np.stack([rng.uniform(5, 109, n), rng.uniform(5, 109, n), rng.uniform(1.0, 3.0, n)], 1)
    pos1 = pos0 + rng.uniform(-0.6, 0.6, pos0.shape) * [1, 1, 0.1]
    mb = sn.LocalMapBuilder(vox, n, rng3)
    L, cells = mb.L, mb.grid_stride
    dim_env = (C.c_int32 * 3)(env.shape[2], env.shape[1], env.shape[0])
    org_c = (C.c_double * 3)(*org)
    d_env, d_p0, d_p1 = dmalloc(env.nbytes), dmalloc(n * 24), dmalloc(n * 24)
    d_g0, d_g1, d_o0, d_o1, d_have = dmalloc(n * cells), dmalloc(n * cells), dmalloc(n * 24), dmalloc(n * 24), dmalloc(n)
    flush_bytes = 512 << 20
    d_flush = dmalloc(flush_bytes)
    h2d(d_env, env), h2d(d_p0, pos0), h2d(d_p1, pos1), h2d(d_have, np.ones(n, np.uint8))
    stream = C.c_void_p()
    ck(rt.cudaStreamCreate(C.byref(stream)))

    def launch(first):
        rc = L.hdsm_sense_batch_device(mb.h, C.c_int(n), d_env, dim_env, org_c, d_p0 if first else d_p1, None,
                                       None if first else d_g0, None if first else d_o0, None if first else d_have,
                                       d_g0 if first else d_g1, d_o0 if first else d_o1, stream)
        if rc != 0:
            raise RuntimeError(L.hdsm_sense_last_error(mb.h).decode())

    def timed(first):
        a, b = C.c_void_p(), C.c_void_p()
        ck(rt.cudaEventCreate(C.byref(a))), ck(rt.cudaEventCreate(C.byref(b)))
        ck(rt.cudaMemsetAsync(d_flush, C.c_int(1), C.c_size_t(flush_bytes), stream))
        ck(rt.cudaEventRecord(a, stream))
        launch(first)
        ck(rt.cudaEventRecord(b, stream))
        ck(rt.cudaStreamSynchronize(stream))
        ms = C.c_float()
        ck(rt.cudaEventElapsedTime(C.byref(ms), a, b))
        return float(ms.value)

    first_ms = timed(True)
    for _ in range(3):
        launch(False)
    ck(rt.cudaStreamSynchronize(stream))
    times = [timed(False) for _ in range(steps)]
    ms = float(np.mean(times))
    got0, got1, goto1 = np.empty((m, cells), np.int8), np.empty((m, cells), np.int8), np.empty((m, 3))
    d2h(got0, d_g0), d2h(got1, d_g1), d2h(goto1, d_o1)
    out = {"kernel_form": "keys" if os.environ.get("HDSM_SENSE_BITS", "")[:1] == "0" else "bits", "n_agents": n, "steps": steps, "kernel_ms": ms, "kernel_ms_min": float(np.min(times)), "first_update_ms": first_ms,
           "agents_per_s": n / (ms * 1e-3), "algorithmic_bytes_per_agent": 2.0 * cells,
           "achieved_gbs": 2.0 * cells * n / (ms * 1e-3) / 1e9, "gpu_launches": mb.launch_count,
           "env_dims": [int(v) for v in dim_env], "rays_per_agent": 2 * (66 * 66 + 2 * 66 * 20)}
    if m > 0:
        from oracle import sensing as osn
        want0, wo0 = osn.c_update(env, org, pos0[:m], vox, rng3)
        t0 = time.perf_counter()
        want1, wo1 = osn.c_update(env, org, pos1[:m], vox, rng3, old_grids=want0, old_origin=wo0)
        t_cpu = time.perf_counter() - t0
        out["byte_exact_vs_cpu_port"] = bool(np.array_equal(got0.reshape(want0.shape), want0) and
                                             np.array_equal(got1.reshape(want1.shape), want1) and np.array_equal(goto1, wo1))
        out["cpu_baseline"] = {"value": m / t_cpu, "unit": "agents/s", "cores": osn.max_threads(), "kind": "port",
                               "sample": f"the first {m} agents once (steady-state update), C restatement on all host threads"}
    mb.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
