"""Layout build time: host builder vs GPU builder on BASELINE config 2 (or --scale R-MAT).  Prints one JSON line."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spmv-fpga_b200"))
import numpy as np
import spmvb

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="laplacian")
ap.add_argument("--scale", type=int, default=22)
ap.add_argument("--cu", type=int, default=1)
ap.add_argument("--cdb", type=int, default=0)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--check", type=int, default=1)
a = ap.parse_args()
if a.workload == "laplacian":
    A = spmvb.Csr.laplacian2d(2048, 2048)
elif a.workload == "rmat":
    A = spmvb.Csr.rmat(a.scale, 16, seed=3)
else:
    A = spmvb.Csr.uniform(1 << a.scale, 1 << a.scale, 16, seed=3)
rp, ci, va = A.row_ptr, A.col_ind, A.values
t0 = time.perf_counter(); host = spmvb.Layout.from_csr(A, a.cu, 1, a.cdb); t_host = (time.perf_counter() - t0) * 1e3
t0 = time.perf_counter(); e0 = spmvb.Engine(host, 0); t_up = (time.perf_counter() - t0) * 1e3
e0.free()
runs = []
import torch
d_rp = torch.from_numpy(np.ascontiguousarray(rp, np.uint64).view(np.int64)).cuda()
d_ci = torch.from_numpy(np.ascontiguousarray(ci, np.uint32).view(np.int32)).cuda()
d_va = torch.from_numpy(np.ascontiguousarray(va, np.float64)).cuda()
torch.cuda.synchronize()
for rep in range(a.reps):
    lay, eng = spmvb.Engine.from_csr(A.rows, A.cols, rp, ci, va, a.cu, 1, True, a.cdb)
    ms = eng.build_ms()
    if rep == 0 and a.check:
        eng.fetch_layout(); assert host.difference(lay) == ""
    eng.free(); lay.free()
    lay, eng = spmvb.Engine.from_csr(A.rows, A.cols, d_rp.data_ptr(), d_ci.data_ptr(), d_va.data_ptr(), a.cu, 1, True, a.cdb, on_device=True)
    ms_dev = eng.build_ms()
    eng.free(); lay.free()
    runs.append({"host_csr": ms, "device_csr": ms_dev})
print(json.dumps({"workload": a.workload, "rows": A.rows, "nnz": A.nnz, "cu": a.cu, "host_builder_ms": t_host,
                  "host_engine_create_ms": t_up, "gpu_builder": runs}))
