"""Row partitions of the R-MAT matrix for 8 GPUs, evaluated on ONE GPU: every shard's SpMV step (clear rows + kernel) is
timed by itself, the slowest shard is what an 8-GPU iteration waits for.  Partitions: equal expected cost
non-zeros + w x (row, block) pairs for several w (w = 0: the non-zero balance of round 1; see bench.rmat_row_bounds).
    python scripts/exp_partition.py <scale> <f32|f64> <parts> w0 w1 ...        JSON lines on stdout"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "spmv-fpga_b200"))
import spmvb  # noqa: E402
from bench import rmat_row_bounds  # noqa: E402

scale, dtype, parts = int(sys.argv[1]), sys.argv[2], int(sys.argv[3])
weights = [float(w) for w in sys.argv[4:]] or [0.0]
isd = dtype == "f64"
vt = np.float64 if isd else np.float32
n = 1 << scale
x = np.random.default_rng(1).random(n).astype(vt)
for w in weights:
    bounds = rmat_row_bounds(scale, parts, pair_weight=w)
    shards = []
    for k in range(parts):
        rb, re = bounds[k], bounds[k + 1]
        t0 = time.time()
        A = spmvb.Csr.rmat(scale, 16, 0.57, 0.19, 0.19, 1, rb, re, isd)
        lay = spmvb.Layout.from_csr(A)
        pairs = int(lay.pairs)
        eng = spmvb.Engine(lay, 0)
        lay.free()
        eng.set_x(x)
        eng.enqueue_steps(3); eng.collect_steps()
        eng.enqueue_steps(20, False, inner_events=False)
        total, _ = eng.collect_steps()
        shards.append(dict(rank=k, rows=re - rb, nnz=int(A.nnz), api_pairs=pairs, variant=int(eng.variant), step_ms=total / 20,
                           setup_s=time.time() - t0))
        eng.free()
        print("w=%g shard %d rows %d nnz %d pairs %d variant %d: %.4f ms" % (w, k, re - rb, A.nnz, pairs, shards[-1]["variant"],
                                                                               shards[-1]["step_ms"]), file=sys.stderr, flush=True)
    line = dict(scale=scale, dtype=dtype, parts=parts, pair_weight=w, bounds=bounds, shards=shards,
                slowest_ms=max(s["step_ms"] for s in shards), mean_ms=float(np.mean([s["step_ms"] for s in shards])))
    print(json.dumps(line), flush=True)
    print("w=%g: slowest %.4f ms, mean %.4f ms" % (w, line["slowest_ms"], line["mean_ms"]), file=sys.stderr, flush=True)
