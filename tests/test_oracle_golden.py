"""CPU suite: pins oracle/ (the restatement used as the checker on the GPU box) against fixtures produced by the
UNMODIFIED reference (tests/golden/make_golden.py -> reference_tiny.npz)."""
import os

import numpy as np
import pytest
import torch

from mars5_tts_b200 import synth, weights
from oracle import ar_oracle, nar_oracle
from tests.golden.inputs import make_inputs

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_tiny.npz"))
torch.set_grad_enabled(False)


@pytest.fixture(scope="module")
def env():
    inp = make_inputs()
    size = inp["size"]
    ar_sd, nar_sd = synth.make_ar_state(size), synth.make_nar_state(size)
    cfg = weights.dims_from_state(ar_sd, nar_sd, None, inp["n_text"])
    return inp, ar_sd, nar_sd, cfg


def test_seeded_inputs_match_fixture(env):
    inp = env[0]
    for k in GOLD.files:
        if k.startswith("chk_"):
            assert abs(float(inp[k[4:]].double().sum()) - float(GOLD[k])) < 1e-6 * max(1.0, abs(float(GOLD[k]))), k
    np.testing.assert_array_equal(inp["ar_prompt"].numpy(), GOLD["ar_prompt"])
    np.testing.assert_array_equal(inp["nar_x"].numpy(), GOLD["nar_x"])


def test_codeclm_forward_logits(env):
    inp, ar_sd, _, cfg = env
    lg = ar_oracle.codeclm_forward(ar_sd, cfg, inp["ar_prompt"], inp["ar_spk"])
    assert np.abs(lg.numpy() - GOLD["ar_logits"]).max() < 2e-5


SCFG = dict(temperature=0.7, top_k=200, top_p=0.2, alpha_frequency=3, alpha_presence=0.4, penalty_window=80,
            eos_penalty_decay=0.5, eos_penalty_factor=1)
SCFG_WIDE = dict(SCFG, temperature=1.0, top_k=50, top_p=0.95, penalty_window=4)


@pytest.mark.parametrize("key,scfg", [("ar_gen_cache", SCFG), ("ar_gen_nocache", SCFG), ("ar_gen_wide", SCFG_WIDE)])
def test_ar_generate_token_exact(env, key, scfg):
    inp, ar_sd, _, cfg = env
    seq, hit = ar_oracle.ar_generate(ar_sd, cfg, inp["ar_prompt"], inp["ar_spk"], scfg, inp["ar_noise"], inp["ar_max_len"],
                                     7, inp["eos"])
    np.testing.assert_array_equal(seq.numpy(), GOLD[key])
    assert hit == (len(GOLD[key]) >= inp["ar_max_len"] - 1)


def test_logit_warpers(env):
    inp = env[0]
    n_text, eos = inp["n_text"], inp["eos"]
    for b in range(3):
        # the reference helpers do not mask text ids; emulate by comparing only ids >= n_text-1 of the full chain
        sc = dict(SCFG, penalty_window=20)
        lp = ar_oracle.warp_logits(inp["smp_logits"][b], inp["smp_prev"][b].tolist(), sc, 1, eos, None)
        ref_rep = torch.from_numpy(GOLD["smp_rep"][b])
        # rebuild the rest of the chain from the reference's penalised logits
        z = ref_rep.clone()
        z[eos] -= 1 * (max(9 - 5, 1) ** 0.5)
        np.testing.assert_allclose(z.numpy(), GOLD["smp_eos"][b], rtol=0, atol=1e-6)
        kp = torch.from_numpy(GOLD["smp_kp"][b])
        # same surviving set and same normalised log-probs when the oracle runs the whole chain with n_gen such that
        # the EOS penalty matches (n_gen = 30 > est would skip it) -> compare the pure top-k/top-p stage instead
        sc2 = dict(SCFG, alpha_frequency=0, alpha_presence=0)
        lp2 = ar_oracle.warp_logits(torch.from_numpy(GOLD["smp_eos"][b]), [], sc2, 1, eos, None)
        assert torch.equal(torch.isfinite(lp2), torch.isfinite(kp))
        assert (lp2[torch.isfinite(kp)] - kp[torch.isfinite(kp)].log_softmax(-1)).abs().max() < 1e-5
        sc3 = dict(sc2, temperature=0.9, top_k=40, top_p=0.9)
        kp2 = torch.from_numpy(GOLD["smp_kp2"][b])
        lp3 = ar_oracle.warp_logits(torch.from_numpy(GOLD["smp_eos"][b]), [], sc3, 1, eos, None)
        assert torch.equal(torch.isfinite(lp3), torch.isfinite(kp2))
        # frequency / presence penalty stage
        prev = inp["smp_prev"][b].tolist()
        z0 = inp["smp_logits"][b].clone()
        pv = torch.tensor(prev[-20:])
        vals, cnts = pv.unique(return_counts=True)
        c = torch.zeros_like(z0, dtype=torch.long)
        c[vals] = cnts
        np.testing.assert_allclose((z0 - c * 3 - (c > 0).float() * 0.4).numpy(), GOLD["smp_rep"][b], atol=1e-6)
        assert lp is not None


def test_nar_forward_logits(env):
    inp, _, nar_sd, cfg = env
    for drop, key in ((False, "nar_logits_cond"), (True, "nar_logits_uncond")):
        lg = nar_oracle.nar_forward(nar_sd, cfg, inp["nar_c_text"], inp["nar_c_codes"], inp["nar_x"], int(GOLD["nar_t"]), drop)
        assert np.abs(lg.numpy() - GOLD[key]).max() < 5e-5, key


@pytest.mark.parametrize("T", [10, 200])
def test_diffusion_tables(T):
    tabs = torch.stack(nar_oracle.diffusion_tables(T)).numpy()
    np.testing.assert_array_equal(tabs, GOLD[f"diff_tables_{T}"])
    np.testing.assert_array_equal(weights.diffusion_schedule(T).numpy(), GOLD[f"diff_tables_{T}"])


@pytest.mark.parametrize("t", [0, 3, 199])
def test_reverse_step_codes(env, t):
    inp = env[0]
    tabs = nar_oracle.diffusion_tables(200)
    xo, _ = nar_oracle.reverse_step(tabs, inp["post_cond"], inp["post_uncond"], inp["post_xt"], inp["post_xk"], inp["post_m"],
                                    t, 3, 0.7, inp["post_u"][0], inp["post_u"][1], 1025)
    np.testing.assert_array_equal(xo.numpy(), GOLD[f"post_out_t{t}"])


@pytest.mark.parametrize("deep", [True, False])
def test_nar_loop_codes(env, deep):
    inp, _, nar_sd, cfg = env
    tag = "deep" if deep else "shallow"
    ncfg = dict(T=int(GOLD["nar_loop_T"]), deep_clone=deep, guidance_w=3, x0_temp=0.7, q0_override_steps=2)
    codes = nar_oracle.nar_infer(nar_sd, cfg, inp["nar_c_text"], inp["nar_c_codes"], inp["nar_loop_x_l0"], ncfg,
                                 inp["nar_loop_x_init"], inp[f"nar_loop_{tag}_u"])
    np.testing.assert_array_equal(codes.numpy(), GOLD[f"nar_loop_{tag}_codes"])


# ------------------------------------------------------------------------------------------------ round-2 fixtures
EXTRA = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_extra.npz"))


@pytest.fixture(scope="module")
def extra():
    from tests.golden.inputs import make_extra_inputs
    return make_extra_inputs()


def test_extra_inputs_match_fixture(extra):
    for tag in ("deep", "shallow"):
        assert abs(extra[f"chk_rp_{tag}_u"] - float(EXTRA[f"chk_rp_{tag}_u"])) < 1e-6 * abs(float(EXTRA[f"chk_rp_{tag}_u"]))
    np.testing.assert_array_equal(extra["typ_logits"].numpy(), EXTRA["typ_logits"])
    assert nar_oracle.get_schedule(6, 2, 2) == EXTRA["rp_times"].tolist() == extra["rp_times"]
    assert nar_oracle.get_schedule(200, 1, 1) == list(range(199, -2, -1))


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_typical_p_matches_reference(extra, name):
    """apply_typical_p chained after top-k / top-p exactly as ar_generate.py:88-97 does (samplers.py:96-122)."""
    temp, k, p, mass = EXTRA[f"typ_{name}_cfg"].tolist()
    sc = dict(SCFG, temperature=temp, top_k=int(k), top_p=p, typical_p=mass, alpha_frequency=0, alpha_presence=0)
    for b in range(4):
        lp = ar_oracle.warp_logits(extra["typ_logits"][b], [], sc, extra["n_text"], -1, None)
        ref = torch.from_numpy(EXTRA[f"typ_{name}"][b])
        assert torch.equal(torch.isfinite(lp), torch.isfinite(ref)), (name, b, int(torch.isfinite(lp).sum()), int(torch.isfinite(ref).sum()))
        fin = torch.isfinite(ref)
        assert 0 < int(fin.sum()) < int(torch.isfinite(extra["typ_logits"][b]).sum())
        assert (lp[fin] - ref[fin].log_softmax(-1)).abs().max() < 2e-5


@pytest.mark.parametrize("deep", [True, False])
def test_nar_infer_repaint_jumps_code_exact(env, extra, deep):
    """perform_simple_inference with get_schedule(T, jump_len=2, jump_n_sample=2) and the unscaled forward step."""
    _, _, nar_sd, cfg = env
    tag = "deep" if deep else "shallow"
    codes = nar_oracle.nar_infer(nar_sd, cfg, extra["rp_c_text"], extra["rp_c_codes"], extra["rp_x_l0"],
                                 dict(T=extra["rp_T"], deep_clone=deep, guidance_w=3, x0_temp=0.7, q0_override_steps=2, jump_len=2,
                                      jump_n_sample=2), extra["rp_x_init"], extra[f"rp_{tag}_u"])
    np.testing.assert_array_equal(codes.numpy(), EXTRA[f"rp_{tag}_codes"])
