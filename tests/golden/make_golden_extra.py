"""Generates tests/golden/reference_extra.npz by running the UNMODIFIED reference (/root/reference/mars5/*) on CPU:
round-2 additions to reference_tiny.npz (same seeded synthetic TINY checkpoints, see make_golden.py).

  * apply_typical_p (samplers.py:96-122) after the top-k / top-p warp, as ar_generate.py:88-97 chains them
  * perform_simple_inference with RePaint jumps (jump_len = jump_n_sample = 2): get_schedule + reverse_diffusion +
    forward_diffusion.  `enable_kevin_scaled_inference=False`: the scaled variant (q_pred_one_timestep_scaled,
    diffuser.py:136-159) broadcasts its (1, S, 1) position ramp against a (1, S, 8, K) tensor and raises for every
    S != 8 in the unmodified reference, so only the unscaled forward step can be pinned.

Run in the build container only:   python tests/golden/make_golden_extra.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.path.insert(0, HERE)

from mars5 import diffuser as ref_diff  # noqa: E402
from mars5.samplers import apply_typical_p, top_k_top_p_filtering  # noqa: E402

from make_golden import build_reference_models  # noqa: E402
from mars5_tts_b200 import synth  # noqa: E402

torch.set_grad_enabled(False)
torch.set_num_threads(8)


def main():
    size = synth.TINY
    lm, nar, ar_sd, nar_sd = build_reference_models(size)
    n_text = 258
    V = n_text + 1025
    out = {}
    g = torch.Generator().manual_seed(4321)

    # ---------------- typical-p
    lg = torch.randn(4, V, generator=g) * 3
    lg[:, : n_text - 1] = float("-inf")
    out["typ_logits"] = lg.numpy()
    for name, (temp, k, p, mass) in {"a": (0.7, 200, 0.9, 0.6), "b": (1.0, 0, 1.0, 0.3), "c": (0.9, 50, 0.95, 0.95)}.items():
        f = top_k_top_p_filtering(lg.clone() / temp, top_k=k, top_p=p)
        out[f"typ_{name}"] = apply_typical_p(f, mass=mass).numpy()
        out[f"typ_{name}_cfg"] = np.array([temp, k, p, mass], dtype=np.float64)

    # ---------------- RePaint jumps, unscaled forward diffusion
    Pf, Tc, N, T = 12, 11, 7, 6
    c_text = torch.randint(0, n_text, (Tc,), generator=g)
    c_codes = torch.randint(0, 1024, (Pf, 8), generator=g)
    x_l0 = torch.randint(0, 1024, (N,), generator=g)
    x_init = torch.randint(0, 1025, (N, 8), generator=g)
    times = ref_diff.get_schedule(T, jump_len=2, jump_n_sample=2)
    out["rp_times"] = np.asarray(times, dtype=np.int64)
    n_draws = sum(2 if (b < a and a > 0) else 1 for a, b in zip(times[:-1], times[1:]))
    real_randint, real_rand_like = torch.randint, torch.rand_like
    for deep in (True, False):
        S_tot = N + (Pf if deep else 0)
        u = torch.rand(n_draws, S_tot, 8, 1025, generator=g)
        st_ = {"i": 0}

        def fake_randint(lo, hi, shape, **kw):
            return x_init[None].clone()

        def fake_rand_like(t_, **kw):
            st_["i"] += 1
            return u[st_["i"] - 1][None].clone().to(t_.dtype)

        torch.randint, torch.rand_like = fake_randint, fake_rand_like
        try:
            diff = ref_diff.MultinomialDiffusion(1025, timesteps=T)
            dsh = ref_diff.DSH(last_greedy=True, x_0_temp=0.7, guidance_w=3, deep_clone=deep, jump_len=2, jump_n_sample=2,
                               q0_override_steps=2, enable_kevin_scaled_inference=False, progress=False)
            _x = x_l0[None, :, None].repeat(1, 1, 8)
            # dtype float32: inference.py:298 passes torch.float16, which only reaches index_to_log_onehot of the forward
            # step (dead at jump 1/1); the fp32 variant is what is pinned here (fp16 would pin torch's half rounding of
            # every elementwise op of the forward step)
            res = ref_diff.perform_simple_inference(nar, (c_text[None], c_codes[None].clone(), torch.tensor([Tc]),
                                                          torch.tensor([Pf]), _x, torch.zeros(1, N, dtype=torch.bool)),
                                                    diff, T, torch.float32, dsh=dsh, retain_quant0=True)
        finally:
            torch.randint, torch.rand_like = real_randint, real_rand_like
        assert st_["i"] == n_draws, (st_["i"], n_draws)
        tag = "deep" if deep else "shallow"
        out[f"rp_{tag}_codes"] = res[0].numpy()
        out[f"chk_rp_{tag}_u"] = np.float64(u.double().sum())
    out.update(rp_c_text=c_text.numpy(), rp_c_codes=c_codes.numpy(), rp_x_l0=x_l0.numpy(), rp_x_init=x_init.numpy(),
               rp_T=np.int64(T), rp_n_draws=np.int64(n_draws))
    path = os.path.join(HERE, "reference_extra.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB;", len(out), "arrays; schedule", times)


if __name__ == "__main__":
    main()
