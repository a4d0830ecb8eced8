"""Replays the seeded input draws of make_golden.py (same generator, same order) so that the big random tensors need
not be stored in the fixture.  Keep in lock-step with make_golden.py; the fixture's chk_* sums guard against drift."""
import torch

from mars5_tts_b200 import synth


def make_inputs():
    size = synth.TINY
    n_text = 258
    V = n_text + 1025
    g = torch.Generator().manual_seed(1234)
    d = {}
    Pf, n_txt, n_sp = 12, 9, 5
    d["ar_spk"] = torch.randint(0, 1024, (Pf, 8), generator=g)
    text_ids = [256] + torch.randint(0, 256, (n_txt,), generator=g).tolist() + [257]
    speech_ids = (torch.randint(0, 1024, (n_sp,), generator=g) + n_text).tolist()
    d["ar_text_ids"], d["ar_prompt"] = text_ids, torch.tensor(text_ids + speech_ids)
    steps = 14
    d["ar_max_len"] = len(d["ar_prompt"]) + 14
    d["ar_noise"] = torch.empty(steps, V).exponential_(1, generator=g)
    d["smp_logits"] = torch.randn(3, V, generator=g) * 3
    d["smp_prev"] = torch.randint(n_text, V, (3, 30), generator=g)
    Tc, S = 11, 19
    d["nar_c_text"] = torch.randint(0, n_text, (Tc,), generator=g)
    d["nar_c_codes"] = torch.randint(0, 1024, (Pf, 8), generator=g)
    d["nar_x"] = torch.randint(0, 1025, (S, 8), generator=g)
    T, N = 6, 7
    d["nar_loop_T"] = T
    d["nar_loop_x_l0"] = torch.randint(0, 1024, (N,), generator=g)
    d["nar_loop_x_init"] = torch.randint(0, 1025, (N, 8), generator=g)
    d["nar_loop_deep_u"] = torch.rand(T, 2, N + Pf, 8, 1025, generator=g)
    d["nar_loop_shallow_u"] = torch.rand(T, 2, N, 8, 1025, generator=g)
    Sx = 9
    d["post_cond"] = torch.randn(Sx, 8, 1025, generator=g) * 2
    d["post_uncond"] = torch.randn(Sx, 8, 1025, generator=g) * 2
    d["post_xt"] = torch.randint(0, 1025, (Sx, 8), generator=g)
    d["post_xk"] = torch.randint(0, 1024, (Sx, 8), generator=g)
    d["post_m"] = torch.rand(Sx, 8, generator=g) < 0.4
    d["post_u"] = torch.rand(2, Sx, 8, 1025, generator=g)
    d["size"], d["n_text"], d["V"], d["eos"] = size, n_text, V, n_text + 1024
    return d


def trim_cases():
    """Seeded mono waveforms for the silence-trim fixtures (make_trim_golden.py / tests/test_trim_cpu.py): speech-like
    bursts between stretches of low-level noise, at 24 kHz.  Returns a list of (name, float32 tensor)."""
    g = torch.Generator().manual_seed(4321)
    cases = []

    def burst(n, segs, floor):
        env = torch.zeros(n)
        for a, b, amp in segs:
            env[a:b] = amp
        return (torch.randn(n, generator=g) * env + floor * torch.randn(n, generator=g)).float()

    cases.append(("lead_tail_silence", burst(48000, [(9000, 30000, 0.3)], 1e-4)))
    cases.append(("two_bursts", burst(60001, [(5000, 12000, 0.2), (40000, 52000, 0.05)], 1e-4)))
    cases.append(("no_silence", burst(24000, [(0, 24000, 0.1)], 0.0)))
    cases.append(("quiet_tail_above_threshold", burst(36000, [(2000, 20000, 0.5), (20000, 36000, 0.03)], 1e-5)))
    cases.append(("quiet_tail_below_threshold", burst(36000, [(2000, 20000, 0.5), (20000, 36000, 0.01)], 1e-5)))
    cases.append(("all_zero", torch.zeros(8000)))
    cases.append(("digital_silence_then_click", burst(20000, [(15000, 15040, 0.9)], 0.0)))
    cases.append(("full_utterance_length", burst(1499 * 320, [(30000, 420000, 0.25)], 3e-4)))
    cases.append(("shortest_legal", burst(1025, [(100, 900, 0.2)], 1e-4)))
    return cases


def make_extra_inputs():
    """Replays make_golden_extra.py's draws (generator seed 4321, same order): typical-p logits and the RePaint case.
    The (n_draws, S, 8, K) uniforms are expanded into the engine's (n_steps, 2, S, 8, K) layout (reverse step: draw 0 =
    unknown sample, draw 1 = known re-noise; forward step: draw 0)."""
    import numpy as np
    n_text = 258
    V = n_text + 1025
    g = torch.Generator().manual_seed(4321)
    d = {}
    lg = torch.randn(4, V, generator=g) * 3
    lg[:, : n_text - 1] = float("-inf")
    d["typ_logits"] = lg
    Pf, Tc, N, T = 12, 11, 7, 6
    d["rp_c_text"] = torch.randint(0, n_text, (Tc,), generator=g)
    d["rp_c_codes"] = torch.randint(0, 1024, (Pf, 8), generator=g)
    d["rp_x_l0"] = torch.randint(0, 1024, (N,), generator=g)
    d["rp_x_init"] = torch.randint(0, 1025, (N, 8), generator=g)
    times = [5, 4, 3, 2, 3, 4, 3, 2, 1, 0, 1, 2, 1, 0, -1]   # get_schedule(6, jump_len=2, jump_n_sample=2)
    n_draws = sum(2 if (b < a and a > 0) else 1 for a, b in zip(times[:-1], times[1:]))
    for tag, S in (("deep", N + Pf), ("shallow", N)):
        u = torch.rand(n_draws, S, 8, 1025, generator=g)
        d[f"chk_rp_{tag}_u"] = float(u.double().sum())
        full = torch.full((len(times) - 1, 2, S, 8, 1025), 0.5)
        i = 0
        for s, (a, b) in enumerate(zip(times[:-1], times[1:])):
            full[s, 0] = u[i]; i += 1
            if b < a and a > 0:
                full[s, 1] = u[i]; i += 1
        assert i == n_draws
        d[f"rp_{tag}_u"] = full
    d["rp_times"], d["rp_T"], d["n_text"], d["V"] = times, T, n_text, V
    return d
