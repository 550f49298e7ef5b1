"""fp64 Gram (A B^T) throughput: the dedicated Gram kernel (lsk_gram.cu: 128 x 64 and 128 x 128 tiles) and the generic tile kernel
(lsk_pairwise.cu, EPI_LINEAR) against cuBLAS DGEMM through torch.matmul; general and symmetric (A A^T) mode.
The tile variant is chosen per process through LSK_GRAM_TILE (2, 4, -1 = generic, unset = automatic): the parent runs one child each.
    python tools/gram_bench.py            -> gpurun_out/gram_bench.json"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SHAPES = [(4096, 4096, 4096), (8192, 8192, 8192), (16384, 16384, 4096), (8192, 8192, 20480), (8192, 8192, 16384)]
SYM_SHAPES = [(16384, 4096), (16384, 20480)]


def child():
    import torch
    from liestationarykernels_b200 import ops

    def timeit(fn, reps=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    with_cublas = os.environ.get("LSK_GRAM_TILE", "") == ""
    res = []
    for (n, m, k) in SHAPES:
        torch.manual_seed(0)
        a = torch.randn(n, k, dtype=torch.float64, device="cuda")
        b = torch.randn(m, k, dtype=torch.float64, device="cuda")
        pa, pb = ops.PanelMatrix.from_dense(a), ops.PanelMatrix.from_dense(b)
        out = torch.empty(n, m, dtype=torch.float64, device="cuda")
        t = timeit(lambda: ops.gram(pa, pb, 1.0, out=out))
        ref = a @ b.T
        err = (out - ref).abs().max().item()
        r = dict(mode="general", n=n, m=m, k=k, ms=t, tflops=2.0 * n * m * k / t / 1e9, max_abs_err_vs_cublas=err)
        if with_cublas:
            tc = timeit(lambda: torch.matmul(a, b.T, out=out))
            r.update(cublas_ms=tc, cublas_tflops=2.0 * n * m * k / tc / 1e9)
        res.append(r)
        del a, b, pa, pb, out, ref
        torch.cuda.empty_cache()
    for (n, k) in SYM_SHAPES:
        torch.manual_seed(0)
        a = torch.randn(n, k, dtype=torch.float64, device="cuda")
        pa = ops.PanelMatrix.from_dense(a)
        out = torch.empty(n, n, dtype=torch.float64, device="cuda")
        t = timeit(lambda: ops.gram(pa, pa, 1.0, out=out), reps=3)
        gen = ops.gram(pa, ops.PanelMatrix.from_dense(a), 1.0)
        r = dict(mode="symmetric", n=n, k=k, ms=t, tflops_computed_half=1.0 * n * n * k / t / 1e9,
                 equals_general_mode=bool(torch.equal(out, gen)), symmetric=bool(torch.equal(out, out.T)))
        res.append(r)
        del a, pa, out, gen
        torch.cuda.empty_cache()
    print(json.dumps(res))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
        sys.exit(0)
    allres = {}
    for tag, val in (("auto", None), ("tile_128x64", "2"), ("tile_128x128", "4"), ("generic_kernel", "-1")):
        env = dict(os.environ)
        env.pop("LSK_GRAM_TILE", None)
        if val is not None:
            env["LSK_GRAM_TILE"] = val
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, capture_output=True, text=True)
        if p.returncode != 0:
            allres[tag] = {"error": p.stderr[-800:]}
            continue
        allres[tag] = json.loads(p.stdout.strip().splitlines()[-1])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(allres, open(os.path.join(ROOT, "gpurun_out", "gram_bench.json"), "w"), indent=1)
    for tag, rows in allres.items():
        print(tag)
        if isinstance(rows, dict):
            print("  ", rows)
            continue
        for r in rows:
            print("  ", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items()})
