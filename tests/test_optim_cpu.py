"""CPU-side checks of the optimizer row: the numpy oracle's code search and update rules,
the exported C symbols, and the host logic of bnb.optim (construction, overrides, state-dict wrapping, loud failure
without a GPU)."""
import ctypes as ct

import numpy as np
import pytest
import torch

import bitsandbytes_b200 as bnb
from bitsandbytes_b200 import cextension
from oracle import optim_ref as R


def _dynamic(signed=True):
    return bnb.functional.create_dynamic_map(signed=signed).numpy().astype(np.float32)


@pytest.mark.parametrize("signed", [True, False])
def test_code_search_returns_a_nearest_entry(signed):
    code = _dynamic(signed)
    rng = np.random.default_rng(0)
    lo = -1.0 if signed else 0.0
    x = np.concatenate([rng.uniform(lo, 1.0, 20000), code, (code[:-1] + code[1:]) / 2, [lo - 0.5, 1.5]]).astype(np.float32)
    got = R.code_search(code, x).astype(np.int64)
    d = np.abs(code[None, :].astype(np.float64) - x[:, None].astype(np.float64))
    best = d.min(1)
    assert np.allclose(d[np.arange(len(x)), got], best, rtol=0, atol=1e-7), "not a nearest entry"
    mids = ((code[:-1] + code[1:]) * np.float32(0.5)).astype(np.float32)
    on_mid = R.code_search(code, mids).astype(np.int64)  # exactly ON a midpoint: the entry the search stopped at
    assert ((on_mid == np.arange(255)) | (on_mid == np.arange(1, 256))).all()


def test_8bit_adam_step_is_the_32bit_step_within_the_quantisation_error():
    rng = np.random.default_rng(1)
    n = 1024
    g = (rng.standard_normal(n) * 0.1).astype(np.float32)
    p = rng.standard_normal(n).astype(np.float32)
    code1, code2 = _dynamic(True), _dynamic(False)
    c1 = np.full(n, int(np.argmin(np.abs(code1))), np.uint8)  # the code of 0.0
    c2 = np.zeros(n, np.uint8)
    a = np.zeros(4, np.float32)
    p8, nc1, nc2, am1, am2 = R.update_8bit_blockwise("adam", "fp32", g, p, c1, c2, code1, code2, a, a.copy(), 1, 1e-3, 0.9, 0.999)
    p32, s1, s2, _ = R.update_32bit("adam", "fp32", g, p, np.zeros(n, np.float32), np.zeros(n, np.float32), 1, 1e-3, 0.9, 0.999)
    np.testing.assert_allclose(p8, p32, rtol=0, atol=2e-6)
    blk = np.arange(n) // 256
    np.testing.assert_allclose(code1[nc1] * am1[blk], s1, atol=0.02 * np.abs(s1).max())
    np.testing.assert_allclose(code2[nc2] * am2[blk], s2, atol=0.02 * np.abs(s2).max())
    assert (np.signbit(code1[nc1]) == np.signbit(s1)).all(), "state1 keeps its sign through quantisation"


def test_the_library_exports_every_optimizer_symbol_of_the_reference_abi():
    dll = ct.CDLL(str(cextension.PACKAGE_DIR / cextension.LIBRARY_NAME))
    names = [f"c{o}32bit_grad_{s}" for o in ("adam", "lion", "ademamix") for s in ("fp32", "fp16", "bf16")]
    names += [f"c{o}32bit_grad_{s}" for o in ("momentum", "rmsprop", "adagrad") for s in ("32", "16")]
    names += [f"c{o}_8bit_blockwise_grad_{s}" for o in ("adam", "momentum", "rmsprop", "adagrad", "lion", "ademamix")
              for s in ("fp32", "fp16", "bf16")]
    names += ["cbnb_b200_optimizer_update_32bit", "cbnb_b200_optimizer_update_8bit_blockwise"]
    missing = [n for n in names if not hasattr(dll, n)]
    assert not missing, missing
    assert set(names) <= set(cextension.EXPORTED_SYMBOLS)


def test_constructors_mirror_the_reference_and_reject_what_it_rejects():
    O = bnb.optim
    p = [torch.nn.Parameter(torch.zeros(8))]
    assert O.Adam8bit(p).args.optim_bits == 8 and O.AdamW(p).defaults["weight_decay"] == 1e-2
    assert O.PagedLion8bit(p).is_paged and O.LAMB(p).args.max_unorm == 1.0 and O.LARS(p, lr=0.1, momentum=0.9).args.max_unorm == 0.02
    assert O.SGD8bit(p, lr=0.1, momentum=0.9).optimizer_name == "momentum" and O.AdEMAMix8bit(p).optimizer_name == "ademamix"
    for bad in (lambda: O.SGD(p, lr=0.1), lambda: O.RMSprop(p, alpha=0), lambda: O.RMSprop(p, centered=True)):
        with pytest.raises(NotImplementedError):
            bad()
    for bad in (lambda: O.Adam8bit(p, amsgrad=True), lambda: O.Adam8bit(p, optim_bits=8), lambda: O.Adagrad(p, lr_decay=0.1),
                lambda: O.Adam(p, lr=-1.0), lambda: O.Adam(p, betas=(1.0, 0.9)), lambda: O.Adagrad8bit(p, optim_bits=32)):
        with pytest.raises(ValueError):
            bad()
    assert O.Adam(p, betas="(0.8, 0.95)").defaults["betas"] == [0.8, 0.95]


def test_per_parameter_overrides_reach_get_config():
    O = bnb.optim
    mng = O.GlobalOptimManager.get_instance()
    mng.initialize()
    a, b = torch.nn.Parameter(torch.zeros(8)), torch.nn.Parameter(torch.zeros(8))
    mng.register_parameters([a, b])
    mng.override_config(b, "optim_bits", 32)
    opt = O.Adam8bit([a, b])
    assert opt.get_config(0, 0, opt.param_groups[0])["optim_bits"] == 8
    assert opt.get_config(0, 1, opt.param_groups[0])["optim_bits"] == 32
    mng.initialize()


def test_state_dict_wraps_the_quantisation_tensors_and_load_restores_them():
    O = bnb.optim
    p = torch.nn.Parameter(torch.zeros(512))
    opt = O.Adam8bit([p])
    opt.state[p] = {"step": 3, "state1": torch.zeros(512, dtype=torch.uint8), "state2": torch.zeros(512, dtype=torch.uint8),
                    "qmap1": torch.zeros(256), "qmap2": torch.zeros(256), "absmax1": torch.ones(2), "absmax2": torch.ones(2)}
    sd = opt.state_dict()
    entry = sd["state"][0]
    assert set(entry) == {"step", O.Adam8bit._FSDP_WRAPPED_QUANT_STATE_KEY}
    assert set(entry[O.Adam8bit._FSDP_WRAPPED_QUANT_STATE_KEY]) == {"state1", "state2", "qmap1", "qmap2", "absmax1", "absmax2"}
    opt2 = O.Adam8bit([torch.nn.Parameter(torch.zeros(512))])
    opt2.load_state_dict(sd, move_to_device=False)
    st = next(iter(opt2.state.values()))
    assert st["step"] == 3 and st["state1"].dtype == torch.uint8 and st["absmax1"].dtype == torch.float32
    assert "state1" in opt.state[p], "state_dict() must not strip the live state"


def test_a_step_on_a_cpu_parameter_fails_loudly():
    p = torch.nn.Parameter(torch.zeros(16))
    p.grad = torch.ones(16)
    with pytest.raises(NotImplementedError):
        bnb.optim.Adam8bit([p]).step()


# ------------------------------------------------------------------------------------------ oracle pinned to the reference
_GOLD = None


def _gold():
    global _GOLD
    if _GOLD is None:
        from pathlib import Path

        _GOLD = np.load(Path(__file__).parent / "golden" / "reference_optim.npz")
    return _GOLD


_HYPER = {"adam": (1e-3, 0.9, 0.999, 0.0, 0.0, 1e-8), "momentum": (1e-2, 0.9, 0.0, 0.0, 0.0, 0.0),
          "rmsprop": (1e-2, 0.99, 0.0, 0.0, 0.0, 1e-8), "adagrad": (1e-2, 0.0, 0.0, 0.0, 0.0, 1e-10),
          "lion": (1e-4, 0.9, 0.99, 0.0, 0.0, 0.0), "ademamix": (1e-3, 0.9, 0.999, 0.9999, 5.0, 1e-8)}


@pytest.mark.parametrize("name", list(_HYPER))
@pytest.mark.parametrize("wd", [0.0, 0.01])
def test_oracle_32bit_update_matches_the_reference_cpu_kernels(name, wd):
    """oracle/optim_ref.update_32bit against outputs of the reference's own CPU kernel (tests/golden/make_golden_optim.py):
    three steps, fp32.  Two known differences bound the tolerance: the CUDA kernels (and the oracle) form 1 - beta in
    fp32 (1 - 0.999f = 0.00100005, 1 - 0.9999f = 0.000100017), the CPU kernel passes the doubles to PyTorch: up to 1.7e-4
    relative on a state;
    and with weight decay the CUDA kernel decays AFTER the update, the CPU kernel before it (lr * wd on the update)."""
    gold = _gold()
    lr, b1, b2, b3, alpha, eps = _HYPER[name]
    tag = f"{name}_wd{int(wd > 0)}"
    n = int(gold["n"])
    p = gold[f"{tag}_p0"].copy()
    s1 = np.zeros((2, n) if name == "ademamix" else (n,), np.float32)
    s2 = np.zeros(n, np.float32) if name in ("adam", "ademamix") else None
    for step in (1, 2, 3):
        p_prev = p
        p, s1, s2, _ = R.update_32bit(name, "fp32", gold[f"{tag}_g"][step - 1], p, s1, s2, step, lr, b1, b2, b3, alpha, eps, wd)
        want = gold[f"{tag}_32_p{step}"]
        upd = np.abs(want - p_prev).max()
        np.testing.assert_allclose(p, want, rtol=0, atol=4e-7 * np.abs(want).max() + (2 * lr * wd + 4e-4) * upd,
                                   err_msg=f"{tag} step {step}: p")
        if not (wd > 0 and name in ("momentum", "rmsprop", "adagrad")):  # (coupled decay enters the state through p)
            np.testing.assert_allclose(s1, gold[f"{tag}_32_s1_{step}"], rtol=3e-4, atol=1e-9, err_msg=f"{tag} step {step}: state1")
            if s2 is not None:
                np.testing.assert_allclose(s2, gold[f"{tag}_32_s2_{step}"], rtol=3e-4, atol=1e-12, err_msg=f"{tag}: state2")


@pytest.mark.parametrize("name", list(_HYPER))
def test_oracle_8bit_blockwise_update_matches_the_reference_cpu_kernel(name):
    """One step from a random mid-training 8-bit state: parameters, new absmax and (almost all) new codes equal the
    reference CPU kernel's -- that kernel re-quantises with an exact nearest-entry search and without the CUDA kernel's
    sign fix, so a code may sit one entry away where a value lies on a midpoint or crosses zero."""
    if name == "ademamix":
        pytest.skip("the golden tensor has 1000 elements: the CUDA kernel (and the oracle) index the slow EMA's absmax at "
                    "(n + i) / 256, which equals the CPU kernel's [2, blocks] layout only for n % 256 == 0")
    gold = _gold()
    lr, b1, b2, b3, alpha, eps = _HYPER[name]
    tag = f"{name}_wd0"
    two = name in ("adam", "ademamix")
    a1 = gold[f"{tag}_8_a1"].reshape(-1)
    got = R.update_8bit_blockwise(name, "fp32", gold[f"{tag}_g"][0], gold[f"{tag}_p0"], gold[f"{tag}_8_c1"],
                                  gold[f"{tag}_8_c2"] if two else None, gold["code1"], gold["code2"], a1,
                                  gold[f"{tag}_8_a2"] if two else None, 2, lr, b1, b2, b3, alpha, eps, 0.0)
    want_p = gold[f"{tag}_8_p"]
    upd = np.abs(want_p - gold[f"{tag}_p0"]).max()
    np.testing.assert_allclose(got[0], want_p, rtol=0, atol=4e-7 * np.abs(want_p).max() + 2e-4 * upd, err_msg=f"{name}: p")
    n = int(gold["n"])
    if name != "ademamix":  # (n = 1000: the reference CUDA indexing of the slow EMA's absmax assumes n % 256 == 0)
        np.testing.assert_allclose(got[3][:len(gold[f"{tag}_8_a1_out"].reshape(-1))], gold[f"{tag}_8_a1_out"].reshape(-1),
                                   rtol=1e-4, err_msg=f"{name}: absmax1")
        c1 = np.asarray(got[1]).reshape(-1)[:n].astype(np.int64)
        w1 = gold[f"{tag}_8_c1_out"].reshape(-1)[:n].astype(np.int64)
        assert np.mean(c1 == w1) > 0.97 and np.abs(c1 - w1).max() <= 1, f"{name}: state1 codes {np.mean(c1 == w1):.4f}"
    if two:
        np.testing.assert_allclose(got[4], gold[f"{tag}_8_a2_out"], rtol=1e-4, err_msg=f"{name}: absmax2")
        c2, w2 = got[2].astype(np.int64), gold[f"{tag}_8_c2_out"].astype(np.int64)
        assert np.mean(c2 == w2) > 0.97 and np.abs(c2 - w2).max() <= 1, f"{name}: state2 codes {np.mean(c2 == w2):.4f}"
