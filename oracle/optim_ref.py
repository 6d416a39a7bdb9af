"""oracle/optim_ref.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy restatement of the reference's optimizer updates: fp32 arithmetic in the
order of the reference kernels, exact division / sqrt / pow where the reference's --use_fast_math build uses
approximations, so GPU results are compared within a stated tolerance (tests/test_gpu_optim.py), while the strict
comparison on the GPU is against the reference CUDA library itself (oracle/_ref/libbitsandbytes_cuda_ref.so).

Parity pinning status of THIS file: pinned by the GPU tests against the reference CUDA library (same inputs through
both libraries); the reference holds no golden vectors for its optimizers (its tests compare with torch.optim within
loose tolerances, reference tests/test_optim.py).

  update_32bit            reference csrc/kernels.cu:605-727 (two states), 806-909 (one state); trust ratio :531-603,
                          :729-804; launch order csrc/ops.cu:80-139
  update_8bit_blockwise   reference csrc/kernels.cu:914-1150 (two states), 1152-1325 (one state)
  code_search             reference csrc/kernels.cu:221-267
"""
from __future__ import annotations

import numpy as np

f32 = np.float32
BLOCK = 256
TWO_STATE = ("adam", "lamb", "ademamix")


def _round_to(x, dtype):
    """fp32 -> storage dtype -> fp32 (dtype: 'fp32' | 'fp16' | 'bf16')."""
    x = np.asarray(x, dtype=f32)
    if dtype == "fp32":
        return x
    if dtype == "fp16":
        return x.astype(np.float16).astype(f32)
    # bf16: round to nearest even on the upper 16 bits
    u = x.view(np.uint32).astype(np.uint64)
    nan = np.isnan(x)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    out = r.view(f32).copy()
    out[nan] = np.nan
    return out


def _sgn(v):
    return np.sign(v).astype(f32)


def update_32bit(name, dtype, g, p, s1, s2, step, lr, beta1, beta2=0.0, beta3=0.0, alpha=0.0, eps=1e-8, weight_decay=0.0,
                 gnorm_scale=1.0, max_unorm=0.0, skip_zeros=False):
    """Returns (p, s1, s2, unorm).  g, p: fp32 arrays holding values of `dtype`; s1 (for ademamix: [2, n]), s2: fp32."""
    lr, beta1, beta2, beta3, alpha, eps, wd, gs = (f32(v) for v in (lr, beta1, beta2, beta3, alpha, eps, weight_decay,
                                                                      gnorm_scale))
    g = _round_to(gs * g.astype(f32), dtype)
    p = p.astype(f32).copy()
    one = f32(1.0)
    unorm = f32(0.0)
    base = "adam" if name == "lamb" else "momentum" if name == "lars" else name
    param_norm = f32(np.linalg.norm(p.astype(np.float64))) if max_unorm > 0 else f32(0.0)

    def trust_sum(a, b):
        if base == "adam":
            c1 = one / (one - f32(beta1) ** f32(step))
            c2 = one / (one - f32(beta2) ** f32(step))
            a = (a * beta1 + (one - beta1) * g) * c1
            b = (b * beta2 + (one - beta2) * (g * g)) * c2
            u = a / (np.sqrt(b) + eps)
            return f32(np.sum((u * u).astype(np.float64)))
        if base == "momentum":
            a = g.copy() if step == 1 else a * beta1 + g
            return f32(np.sum((a * a).astype(np.float64)))
        if base == "lion":
            a = a * beta2 + (one - beta2) * g
            return f32(np.sum(a.astype(np.float64)))
        if base == "rmsprop":
            a = a * beta1 + (one - beta1) * g * g
        else:  # adagrad
            a = a + g * g
        u = g / (np.sqrt(a) + eps)
        return f32(np.sum((u * u).astype(np.float64)))

    def scale_from(unorm, two):
        if max_unorm <= 0:
            return one
        us = f32(np.sqrt(unorm))
        cap = f32(max_unorm) * param_norm + (f32(0.0) if two else eps)
        return cap / us if us > cap else one

    if base in ("adam", "ademamix"):
        s1 = s1.astype(f32).copy()
        s2 = s2.astype(f32).copy()
        if base == "ademamix":
            c1 = one - f32(beta1) ** f32(step)
            c2 = f32(np.sqrt(one - f32(beta2) ** f32(step)))
            m1, m2 = s1[0], s1[1]
            m1 = m1 * beta1 + (one - beta1) * g
            m2 = m2 * beta3 + (one - beta3) * g
            s2 = s2 * beta2 + (one - beta2) * g * g
            p = _round_to(p - lr * (((m1 / c1) + (alpha * m2)) / ((np.sqrt(s2) / c2) + eps)), dtype)
            if wd > 0:
                p = _round_to(p * (one - lr * wd), dtype)
            return p, np.stack([m1, m2]), s2, unorm
        if max_unorm > 0:
            unorm = trust_sum(s1, s2)
        us = scale_from(unorm, True)
        c1 = one - f32(beta1) ** f32(step)
        c2 = f32(np.sqrt(one - f32(beta2) ** f32(step)))
        step_size = -lr * c2 / c1
        act = np.ones_like(g, dtype=bool) if not skip_zeros else g != 0
        a = np.where(act, s1 * beta1 + (one - beta1) * g, s1).astype(f32)
        b = np.where(act, s2 * beta2 + (one - beta2) * (g * g), s2).astype(f32)
        pn = _round_to(p + us * step_size * (a / (np.sqrt(b) + eps * c2)), dtype)
        if wd > 0:
            pn = _round_to(pn * (one - lr * wd), dtype)
        return np.where(act, pn, p).astype(f32), a, b, unorm

    s1 = s1.astype(f32).copy()
    if wd > 0 and base != "lion":
        g = _round_to(g + p * wd, dtype)
    if max_unorm > 0 and base != "lion":
        unorm = trust_sum(s1, None)
    us = scale_from(unorm, False)
    act = np.ones_like(g, dtype=bool) if not skip_zeros else g != 0
    if base == "momentum":
        a = g.copy() if step == 1 else s1 * beta1 + g
        pn = _round_to(p + us * (-lr * a), dtype)
    elif base == "lion":
        pd = _round_to(p * (one - lr * wd), dtype) if wd > 0 else p
        pn = _round_to(pd - us * (lr * _sgn(s1 * beta1 + (one - beta1) * g)), dtype)
        a = s1 * beta2 + (one - beta2) * g
    elif base == "rmsprop":
        a = s1 * beta1 + (one - beta1) * g * g
        pn = _round_to(p - us * (lr * (g / (np.sqrt(a) + eps))), dtype)
    else:
        a = s1 + g * g
        pn = _round_to(p - lr * (g / (np.sqrt(a) + eps)), dtype)
    a = np.where(act, a, s1).astype(f32)
    pn = np.where(act, pn, p).astype(f32)
    if max_unorm > 0 and base == "lion":
        unorm = trust_sum(a, None)
    return pn, a, None, unorm


def code_search(code, x):
    """Index of the nearest entry of the sorted 256-entry code book (7-step search + midpoint rule; a value exactly on a
    midpoint stays with the entry the search stopped at), vectorised over x."""
    code = np.asarray(code, dtype=f32)
    x = np.asarray(x, dtype=f32)
    pivot = np.full(x.shape, 127, dtype=np.int64)
    upper = np.full(x.shape, 255, dtype=np.int64)
    lower = np.zeros(x.shape, dtype=np.int64)
    val = code[pivot]
    i = 64
    while i > 0:
        gt = x > val
        lower = np.where(gt, pivot, lower)
        upper = np.where(gt, upper, pivot)
        pivot = pivot + np.where(gt, i, -i)
        val = code[pivot]
        i >>= 1
    gt = x > val
    mid_u = (code[upper] + val) * f32(0.5)
    mid_l = (code[lower] + val) * f32(0.5)
    return np.where(gt, np.where(x > mid_u, upper, pivot), np.where(x < mid_l, lower, pivot)).astype(np.uint8)


def _quant_signed(code, s, absmax):
    with np.errstate(divide="ignore", invalid="ignore"):
        c = code_search(code, (s / absmax).astype(f32)).astype(np.int64)
    flip = np.signbit(np.asarray(code, dtype=f32)[c]) != np.signbit(s)
    c = np.where(flip, c + np.where(s > 0, 1, -1), c)
    return (c & 0xFF).astype(np.uint8)


def update_8bit_blockwise(name, dtype, g, p, c1, c2, code1, code2, absmax1, absmax2, step, lr, beta1, beta2=0.0, beta3=0.0,
                          alpha=0.0, eps=1e-8, weight_decay=0.0, gnorm_scale=1.0, skip_zeros=False):
    """Returns (p, c1, c2, absmax1, absmax2).  c1 / c2: uint8 state codes (ademamix: c1 is [2, n], absmax1 flat [2*blocks]
    indexed as the reference indexes it), code1 / code2: the 256-entry code books."""
    lr, beta1, beta2, beta3, alpha, eps, wd, gs = (f32(v) for v in (lr, beta1, beta2, beta3, alpha, eps, weight_decay,
                                                                      gnorm_scale))
    one = f32(1.0)
    n = g.size
    nb = -(-n // BLOCK)
    pad = nb * BLOCK - n
    code1 = np.asarray(code1, dtype=f32)
    g_raw = np.concatenate([g.astype(f32), np.zeros(pad, f32)])
    pp = np.concatenate([p.astype(f32), np.zeros(pad, f32)])
    valid = np.arange(nb * BLOCK) < n
    blk = np.arange(nb * BLOCK) // BLOCK
    absmax1 = absmax1.astype(f32).copy()
    two = name in ("adam", "ademamix")

    def padded(c, fill):
        return np.concatenate([c.astype(np.uint8), np.full(pad, fill, np.uint8)])

    if two:
        code2 = np.asarray(code2, dtype=f32)
        absmax2 = absmax2.astype(f32).copy()
        cc1 = padded(c1.reshape(-1)[:n], 128)
        cc2 = padded(c2, 0)
        finite = np.isfinite(g_raw)
        gsc = (g_raw * gs).astype(f32)
        s2 = code2[cc2] * absmax2[blk]
        s2 = s2 * beta2 + (one - beta2) * gsc * gsc
        s1 = code1[cc1] * absmax1[blk]
        s1 = s1 * beta1 + (one - beta1) * gsc
        s1 = np.where(finite, s1, 0).astype(f32)
        s2 = np.where(finite, s2, 0).astype(f32)
        if name == "ademamix":
            cc3 = padded(c1.reshape(-1)[n:2 * n], 128)
            idx3 = (n + np.arange(nb) * BLOCK) // BLOCK  # the reference's (n + i) / 256
            s3 = code1[cc3] * absmax1[idx3][blk]
            s3 = np.where(finite, s3 * beta3 + (one - beta3) * gsc, 0).astype(f32)
            m3 = np.abs(s3).reshape(nb, BLOCK).max(1)
        m1 = np.abs(s1).reshape(nb, BLOCK).max(1)
        m2 = np.abs(s2).reshape(nb, BLOCK).max(1)
        absmax1[:nb] = m1
        absmax2[:nb] = m2
        if name == "ademamix":
            absmax1[idx3] = m3  # (overlaps the first state's last block when n % 256 != 0, as in the reference)
        cr1 = one - f32(beta1) ** f32(step)
        cr2 = f32(np.sqrt(one - f32(beta2) ** f32(step)))
        if name == "ademamix":
            pn = _round_to(pp - lr * (((s1 / cr1) + (alpha * s3)) / ((np.sqrt(s2) / cr2) + eps)), dtype)
        else:
            step_size = -lr * cr2 / cr1
            pn = _round_to(pp + step_size * (s1 / (np.sqrt(s2) + cr2 * eps)), dtype)
        if wd > 0:
            pn = _round_to(pn * (one - lr * wd), dtype)
        pn = np.where(finite, pn, pp).astype(f32)
        nc1 = _quant_signed(code1, s1, m1[blk])
        with np.errstate(divide="ignore", invalid="ignore"):
            nc2 = code_search(code2, (s2 / m2[blk]).astype(f32))
        if name == "ademamix":
            nc3 = _quant_signed(code1, s3, m3[blk])
            nc1 = np.stack([nc1[:n], nc3[:n]])
            return pn[:n], nc1, nc2[:n], absmax1, absmax2
        return pn[:n], nc1[:n], nc2[:n], absmax1, absmax2

    cc1 = padded(c1, 128)
    gsc = (g_raw * gs).astype(f32)
    act = np.ones_like(valid) if not skip_zeros else g_raw != 0
    s1 = (code1[cc1] * absmax1[blk]).astype(f32)
    if wd > 0:
        if name == "lion":
            pp = np.where(act, _round_to(pp * (one - lr * wd), dtype), pp).astype(f32)
        else:
            gsc = (gsc + pp * wd).astype(f32)
    gl = g_raw
    if name == "momentum":
        a = gsc.copy() if step == 1 else s1 * beta1 + gsc
    elif name == "lion":
        gl = _round_to(lr * _sgn(s1 * beta1 + (one - beta1) * gsc), dtype)
        a = s1 * beta2 + (one - beta2) * gsc
    elif name == "rmsprop":
        a = s1 * beta1 + (one - beta1) * (gsc * gsc)
    else:
        a = s1 + gsc * gsc
    a = np.where(act, a, s1).astype(f32)
    m1 = np.abs(a).reshape(nb, BLOCK).max(1)
    absmax1[:nb] = m1
    if name == "momentum":
        pn = _round_to(pp - lr * a, dtype)
    elif name == "lion":
        pn = _round_to(pp - gl, dtype)
    else:  # rmsprop / adagrad: the UNSCALED gradient, as the reference
        pn = _round_to(pp - lr * (g_raw / (np.sqrt(a) + eps)), dtype)
    pn = np.where(act, pn, pp).astype(f32)
    nc1 = _quant_signed(code1, a, m1[blk])
    return pn[:n], nc1[:n], None, absmax1, None
