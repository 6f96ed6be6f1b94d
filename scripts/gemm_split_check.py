"""Accuracy and stall counters of the two operand splits of the CTA-pair GEMM (mdb_gemm_knob SPLIT):
0 = 3xTF32, 1 = TF32 + two BF16 cross terms.  Errors are against a float64 product of the same fp32 inputs,
relative to the rms of the result; every operand-layout combination (K-major / MN-major A and B) is covered,
with ragged M / N / K.      python scripts/gemm_split_check.py [stalls]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import minidiff_b200.backend as B
from minidiff_b200.backend import functions as F
from minidiff_b200.backend._lib import check, lib

SPLIT = 8
check(lib.mdb_gemm_tune(4 | 32))          # force the CTA-pair kernel
rng = np.random.default_rng(7)
worst = {}
# (split, chain length in k-blocks, truncation compensation in 1e-10 per instruction); printed: max/rms error
CASES = tuple((sp, 4, g) for sp in (0, 1) for g in (0, 100, 150, 200, 250, 300, 400))
for (M, K, N) in ((512, 1024, 768), (300, 1000, 520), (1024, 4096, 512)):
    for ta, tb in ((0, 0), (1, 1)):
        if True:
            ah = rng.standard_normal((K, M) if ta else (M, K), dtype=np.float32)
            bh = rng.standard_normal((N, K) if tb else (K, N), dtype=np.float32)
            if K == 1000:                          # wide dynamic range, mixed signs
                ah *= np.exp(rng.uniform(-8, 8, ah.shape)).astype(np.float32)
            a = B.asarray(ah).T if ta else B.asarray(ah)
            b = B.asarray(bh).T if tb else B.asarray(bh)
            ref = (ah.T if ta else ah).astype(np.float64) @ (bh.T if tb else bh).astype(np.float64)
            rms = np.sqrt(np.mean(ref ** 2))
            line = f"M{M} K{K} N{N} A{'mn' if ta else 'k '} B{'mn' if tb else 'k '}:"
            for split, chunk, gain in CASES:
                check(lib.mdb_gemm_knob(SPLIT, split))
                check(lib.mdb_gemm_knob(9, chunk))
                check(lib.mdb_gemm_knob(10, gain))
                out = B.zeros((M, N), dtype=np.float32)
                F._gemm(a, b, out=out)
                got = out.numpy().astype(np.float64)
                err = np.abs(got - ref)
                e_max, e_rms = err.max() / rms, np.sqrt(np.mean(err ** 2)) / rms
                worst[(split, chunk, gain)] = max(worst.get((split, chunk, gain), 0.0), e_max)
                line += f"  s{split}g{gain}: {e_max:.2e}/{e_rms:.2e}"
            f32 = (ah.T if ta else ah) @ (bh.T if tb else bh)
            e32 = np.abs(f32.astype(np.float64) - ref)
            line += f"  | sgemm: {e32.max() / rms:.2e}/{np.sqrt(np.mean(e32 ** 2)) / rms:.2e}"
            print(line, flush=True)
print("worst max-error / rms:", worst)
check(lib.mdb_gemm_knob(SPLIT, -1))
check(lib.mdb_gemm_knob(9, -1))
check(lib.mdb_gemm_knob(10, -1))
if len(sys.argv) > 1 and sys.argv[1] == "stalls":
    # needs MDB_GEMM_TIMING=1 in the environment
    for M, K, N in ((16384, 4096, 4096),):
        a = B.asarray(rng.standard_normal((M, K), dtype=np.float32))
        b = B.asarray(rng.standard_normal((K, N), dtype=np.float32))
        at = B.asarray(rng.standard_normal((K, M), dtype=np.float32)).T
        out = B.zeros((M, N), dtype=np.float32)
        for split in (0, 1):
            check(lib.mdb_gemm_knob(SPLIT, split))
            for nm, aa in (("A k-major, B mn-major", a), ("A mn-major, B mn-major", at)):
                for _ in range(2):
                    print(f"--- split {split} {nm} {M}x{K}x{N}", file=sys.stderr, flush=True)
                    F._gemm(aa, b, out=out)
                B.synchronize()
    check(lib.mdb_gemm_knob(SPLIT, 0))
