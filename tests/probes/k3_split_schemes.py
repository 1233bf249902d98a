"""Error budget of operand-split schemes for the batched kernel K3 (CPU simulation, float64 accumulation).

K3 computes sim = x . t for fp32 operands on the tensor cores by splitting them into narrow terms.  The kernel in the
tree uses bf16x2 with three MMAs per product (x1*t1 + x2*t1 + x1*t2).  This probe measures, on VQSYN-1 rows, what
cheaper splits would cost in score error against the float64 oracle — the north-star tolerance is 1e-5 relative on the
score — so that the next kernel experiment starts from numbers:

  bf16 x3      x1*t1 + x2*t1 + x1*t2            (3 MMAs, kind::f16)      the current kernel's arithmetic
  bf16 x2      x1*t1 + x2*t1                    (2 MMAs)                 t rounded to bf16 once
  fp16 x1      x1*t1                            (1 MMA)                  operands scaled by powers of two into fp16 range
  fp16 x2      x1*t1 + x2*t1                    (2 MMAs)                 full x, t rounded to fp16 once
  fp16 x2+e4m3 x1*t1 + x2*t1 + q8(x)*q8(t-t1)   (2 MMAs + 1 fp8 MMA at twice the rate = 2.5)   the t residual through e4m3
  fp16 x3      x1*t1 + x2*t1 + x1*t2            (3 MMAs)
  fp16 x1+2e4m3  x1*t1 + q8(x-x1)*q8(t) + q8(x)*q8(t-t1)   (1 MMA + one fp8 MMA over a doubled K = 2.0)

Only operand rounding is modelled (products exact, float64 sums): the tensor core's truncating fp32 accumulation adds
the bias K3 already handles with its two-level accumulation.

  python tests/probes/k3_split_schemes.py [n_clips] [n_queries]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import scoring as sc  # noqa: E402  (checker)
from oracle import synth  # noqa: E402

SEED, REF = 20261018, 18120


def to_bf16(a):
    """round-to-nearest-even to 8 significant bits, returned as float32"""
    u = np.ascontiguousarray(a, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def to_fp16(a):
    return np.asarray(a, np.float32).astype(np.float16).astype(np.float32)


def to_e4m3(a):
    """fp8 e4m3 as the tensor core reads it: 4 significant bits, largest finite value 448, smallest normal 2^-6,
    subnormals in steps of 2^-9 (smaller magnitudes flush to zero); round to nearest."""
    a = np.asarray(a, np.float64)
    m, e = np.frexp(a)                                          # a = m * 2^e, 0.5 <= |m| < 1
    step = np.ldexp(1.0, np.maximum(e - 4, -9))                 # spacing: 4 significant bits, never finer than 2^-9
    return np.clip(np.round(a / step) * step, -448.0, 448.0)


def pow2_scale(a, top):
    """power of two that brings max|a| just below `top` (exact to undo)"""
    mx = float(np.max(np.abs(a)))
    return 2.0 ** np.floor(np.log2(top / mx)) if mx > 0 else 1.0


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    nq = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    X32 = synth.database(SEED, n)                                # [n, 2, 1024] float32
    X = X32.astype(np.float64)
    weights = (1.0, 1.5)
    rows = []
    for q in range(nq):
        ref = synth.rows(SEED, [REF + 37 * q]).astype(np.float64)[0][:, None, :]
        T = sc.scale_target(ref)[:, 0, :]                        # [2, 1024] float64
        T32 = T.astype(np.float32)
        exact = sc.scores(np.einsum("nsd,sd->ns", X, T32.astype(np.float64)), weights)

        def score_of(sims):
            return sc.scores(sims, weights)

        def dots(xs, ts):
            return sum(np.einsum("nsd,sd->ns", a.astype(np.float64), b.astype(np.float64)) for a, b in zip(xs, ts))

        out = {}
        x1, t1 = to_bf16(X32), to_bf16(T32)
        x2, t2 = to_bf16(X32 - x1), to_bf16(T32 - t1)
        out["bf16 x3"] = score_of(dots([x1, x2, x1], [t1, t1, t2]))
        out["bf16 x2"] = score_of(dots([x1, x2], [t1, t1]))
        # fp16: scale each operand by a power of two so that its largest value sits near 2^14 (no overflow, few subnormals)
        sx = pow2_scale(X32, 2.0 ** 14)
        st = np.array([pow2_scale(T32[s], 2.0 ** 14) for s in range(2)])[:, None]
        Xs, Ts = (X32 * sx).astype(np.float32), (T32 * st).astype(np.float32)
        h1, g1 = to_fp16(Xs), to_fp16(Ts)
        h2, g2 = to_fp16(Xs - h1), to_fp16(Ts - g1)
        un = 1.0 / (sx * st.reshape(1, 2))
        out["fp16 x1"] = score_of(dots([h1], [g1]) * un)
        out["fp16 x2"] = score_of(dots([h1, h2], [g1, g1]) * un)
        e = (Ts - g1).astype(np.float64)                         # the t residual, per stream scaled into e4m3's range
        se = np.array([pow2_scale(e[s], 256.0) for s in range(2)])[:, None]
        corr = np.einsum("nsd,sd->ns", to_e4m3(Xs.astype(np.float64) * pow2_scale(Xs, 256.0)), to_e4m3(e * se)) / (pow2_scale(Xs, 256.0) * se.reshape(1, 2))
        out["fp16 x2+e4m3"] = score_of((dots([h1, h2], [g1, g1]) + corr) * un)
        out["fp16 x3"] = score_of(dots([h1, h2, h1], [g1, g1, g2]) * un)
        # one fp16 MMA + both residual terms through e4m3 (one fp8 MMA over a doubled K = 1 bf16-MMA-equivalent): 2.0
        dx = (Xs - h1).astype(np.float64)
        sdx = pow2_scale(dx, 256.0)
        sg = np.array([pow2_scale(Ts[s], 256.0) for s in range(2)])[:, None]
        sxx = pow2_scale(Xs, 256.0)
        c1 = np.einsum("nsd,sd->ns", to_e4m3(dx * sdx), to_e4m3(Ts.astype(np.float64) * sg)) / (sdx * sg.reshape(1, 2))
        c2 = np.einsum("nsd,sd->ns", to_e4m3(Xs.astype(np.float64) * sxx), to_e4m3(e * se)) / (sxx * se.reshape(1, 2))
        out["fp16 x1+2e4m3"] = score_of((dots([h1], [g1]) + c1 + c2) * un)
        # the same with scales a kernel can use with ONE accumulator: the three products carry the same power of two.
        # e4m3 operands top out at 128 (x, t) and, 2^12 finer, at <= 256 (x - x1, t - t1); fp16 operands sit 2^6 above
        # the e4m3 ones: (x 2^6)(t 2^6) = (dx 2^12)(t) = (x)(dt 2^12).  Per store: one scale for x; per (query, stream): one for t.
        bx = pow2_scale(X32, 128.0)
        bt = np.array([pow2_scale(T32[s], 128.0) for s in range(2)])[:, None]
        xa, ta = X32.astype(np.float64) * bx, T32.astype(np.float64) * bt            # exact (powers of two)
        xh, th = to_fp16(xa * 64.0).astype(np.float64), to_fp16(ta * 64.0).astype(np.float64)
        dxa, dta = (xa * 64.0 - xh) * 64.0, (ta * 64.0 - th) * 64.0                 # residuals on the 2^12 scale
        acc = (np.einsum("nsd,sd->ns", xh, th) + np.einsum("nsd,sd->ns", to_e4m3(dxa), to_e4m3(ta))
               + np.einsum("nsd,sd->ns", to_e4m3(xa), to_e4m3(dta)))
        out["same, 1 accum"] = score_of(acc / (4096.0 * bx * bt.reshape(1, 2)))
        for k, v in out.items():
            rel = np.abs(v - exact) / np.maximum(np.abs(exact), 0.05)
            rows.append((k, float(rel.max()), float(np.sqrt(np.mean(rel ** 2))), float(np.mean(v - exact))))
    print("%d clips x %d queries; score error relative to float64 (tolerance 1e-5)" % (n, nq))
    print("%-14s %-12s %-12s %-12s" % ("scheme", "max rel", "rms rel", "mean (bias)"))
    for k in dict.fromkeys(r[0] for r in rows):
        sel = [r for r in rows if r[0] == k]
        print("%-14s %-12.3e %-12.3e %-+12.3e" % (k, max(r[1] for r in sel), np.sqrt(np.mean([r[2] ** 2 for r in sel])),
                                                  np.mean([r[3] for r in sel])))


if __name__ == "__main__":
    main()
