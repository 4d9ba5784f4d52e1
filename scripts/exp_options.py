"""A/B of engine options on ONE generated matrix: python scripts/exp_options.py <workload> <scale> <dtype> "k=v,k=v" ...
Every option set builds its own layout + engine (host builder), times `steps` device SpMVs (CUDA events, kernel only) and
checks the result against the oracle once.  Prints one line per option set; JSON lines go to stdout."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spmv-fpga_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_api as oa  # noqa: E402
import spmvb  # noqa: E402

workload, scale, dtype = sys.argv[1], int(sys.argv[2]), sys.argv[3]
sets = sys.argv[4:] or [""]
isd = dtype == "f64"
vt = np.float64 if isd else np.float32
t0 = time.time()
if workload == "uniform":
    A = spmvb.Csr.uniform(1 << scale, 1 << scale, 16, 1, 0, 0, isd)
elif workload == "rmat":
    A = spmvb.Csr.rmat(scale, 16, 0.57, 0.19, 0.19, 1, 0, 0, isd)
else:
    A = spmvb.Csr.laplacian2d(2048, 1 << (scale - 11), 0, 0, isd)
print("generated %s scale %d: %d nnz in %.1f s" % (workload, scale, A.nnz, time.time() - t0), file=sys.stderr)
x = np.random.default_rng(1).random(A.cols).astype(vt)
O = oa.OracleLib()
gold, _ = O.spmv_gold_omp(A.rows, A.row_ptr, A.col_ind, A.values, x, isd)
bound = O.abs_ax(A.rows, A.row_ptr, A.col_ind, A.values, x, isd) * (1e-12 if isd else 1e-5) + 1e-300
vb = 8 if isd else 4
for s in sets:
    opts = {kv.split("=")[0]: int(kv.split("=")[1]) for kv in s.split(",") if kv and not kv.startswith("variant")}
    variant = [int(kv.split("=")[1]) for kv in s.split(",") if kv.startswith("variant")]
    with spmvb.options(**opts):
        t0 = time.time()
        lay = spmvb.Layout.from_csr(A)
        t_lay = time.time() - t0
        eng = spmvb.Engine(lay, 0, variant[0] if variant else 0)
        lay.free()
        eng.set_x(x)
        eng.enqueue_steps(3); eng.collect_steps()
        eng.enqueue_steps(10, False, inner_events=True)
        total, ker = eng.collect_steps()
        eng.spmv_dev()
        y = eng.get_y()
        err = float(np.max(np.abs(y.astype(np.float64) - gold.astype(np.float64)) / bound))
        alg = eng.algorithmic_bytes
        line = dict(workload=workload, scale=scale, dtype=dtype, options=s, variant=eng.variant, kernel_ms=float(np.mean(ker)),
                kernel_ms_min=float(np.min(ker)), step_ms=total / 10, gbs=alg / float(np.mean(ker)) / 1e6,
                frac=alg / float(np.mean(ker)) / 1e6 / 6547.2, err_over_tol=err, layout_s=t_lay, device_layout=eng.device_layout)
        print(json.dumps(line), flush=True)
        print("%-40s v%d kernel %.4f ms (min %.4f) step %.4f  %.0f GB/s frac %.3f err %.2g  %s" %
          (s, eng.variant, line["kernel_ms"], line["kernel_ms_min"], line["step_ms"], line["gbs"], line["frac"], err,
           {k: line["device_layout"][k] for k in ("cu", "cdb", "cu_major", "e2e_tiles", "tall")}), file=sys.stderr, flush=True)
        assert err <= 1.0 or "diag_flags" in s, "wrong result"
        eng.free()
