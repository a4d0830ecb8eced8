#!/usr/bin/env python
"""bench.py -- synthesized audio seconds per second (deep-clone, batch 32) for the MARS5 AR -> NAR -> vocoder hot path.

One "step" = one pass of the hot path over one batch of B synthetic utterances (BASELINE.json configs[2]):
6 s reference clip (450 Encodec frames), 35-token reference transcript, 100-token target text, generation length forced
to N = 15 frames per text token = 1500 frames (random weights never emit EOS), T = 200 reverse steps with classifier-free
guidance.  value  = audio seconds / second with inputs resident in HBM; e2e = the same through the host-buffer C ABI.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c3|c5]
                    [--mode fast|mixed|mixed8|mixed8k|precise] [--also MODE,MODE]

N > 1: launched under torch.distributed.run, one rank per GPU; utterances shard by batch (weak scaling), weights are
repacked on rank 0 and broadcast over NCCL, finished waveforms are all-gathered.  --impl reference times the reference's
CPU path on the host cores on a bounded sample of the same workload (rank 0 only): the UNMODIFIED reference modules when
/root/reference is importable (build container), otherwise the oracle port (GPU box).

Wall budget: the driver kills a run after 870 s.  Warm-up passes are the full workload at T = 8 reverse steps (they
touch every kernel and shape and size the workspace) except the last, which is a complete step and calibrates the step
time; if K steps do not fit into M5_BENCH_BUDGET_S (default 780 s from process start) fewer steps are timed and the line
says so (`steps` = measured, `steps_requested` = K).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

T_PROC0 = time.perf_counter()

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "audio sec/s synthesized (deep-clone, batch 32)"
UNIT = "audio_s/s"
BUDGET_S = float(os.environ.get("M5_BENCH_BUDGET_S", "780"))
MODES = {"fast": 0, "precise": 1, "mixed": 2, "mixed8": 3, "mixed8k": 4}


def elapsed():
    return time.perf_counter() - T_PROC0


# ------------------------------------------------------------------------------------------------ workload
def make_workload(size, B, seed, n_text_tok=100, n_ref_tok=35, Pf=450, frames_per_tok=15, utt_base=0, mixed=False):
    """Synthetic inputs of SURVEY.md 8(d): token ids stand in for tokenised text, codes for the Encodec-encoded clip.
    mixed: per-utterance text length ~ U{20..120} tokens (BASELINE configs[3])."""
    g = torch.Generator().manual_seed(seed)
    n_text = size["n_text"]
    wl = dict(B=B, Pf=Pf, n_text=n_text, prompts=[], spk=[], text=[], n_phones=[], utt=[], N_b=[])
    for b in range(B):
        ntok = int(torch.randint(20, 121, (1,), generator=g)) if mixed else n_text_tok
        sot, eot = n_text - 2, n_text - 1
        text_full = [sot] + torch.randint(0, n_text - 2, (n_ref_tok + ntok,), generator=g).tolist() + [eot]
        spk = torch.randint(0, 1024, (Pf, 8), generator=g).numpy().astype(np.int32)
        speech_prompt = (spk[:, 0] + n_text).tolist()  # ratio-1 synthetic speech tokenizer: one token per frame
        wl["prompts"].append(np.asarray(text_full + speech_prompt, dtype=np.int32))
        wl["spk"].append(spk)
        wl["text"].append(np.asarray(text_full, dtype=np.int32))
        wl["n_phones"].append(5 * ntok)
        wl["utt"].append(utt_base + b)
        wl["N_b"].append(frames_per_tok * ntok)
    wl["N"] = max(wl["N_b"])
    wl["first_codec_idx"] = len(wl["text"][0]) + 1          # inference.py:256
    wl["max_len"] = max(len(p) for p in wl["prompts"]) + wl["N"] + 2   # generate_max_len_override so the sequence fits
    wl["audio_s"] = sum(n - 1 for n in wl["N_b"]) / 75.0     # frames vocoded per utterance after both crops
    return wl


def run_step(eng, icfg, wl, T, mode, seed=0):
    """AR -> (host glue of inference.py:272-283) -> NAR -> crop -> vocoder through the host-buffer API."""
    eos = eng.dims["ar_vocab"] - 1
    acfg = eng.make_ar_cfg(icfg, wl["max_len"], eos, force_len=wl["N"], sync_every=64)
    ids, _, _ = eng.ar_generate(wl["prompts"], wl["spk"], wl["n_phones"], acfg, seed=seed, utt_ids=wl["utt"])
    l0 = [((np.clip(s.astype(np.int64) - wl["n_text"], 0, None))[wl["first_codec_idx"]:] % 1024).astype(np.int32) for s in ids]
    ncfg = eng.make_nar_cfg(icfg, T=T, precise=mode)
    codes = eng.nar_infer(wl["text"], wl["spk"], l0, ncfg, seed=seed, utt_ids=wl["utt"])
    outs = [c[wl["Pf"]:] for c in codes]                     # second crop, inference.py:300-301
    return eng.vocode(outs, bandwidth_id=1)


def to_device(wl, dev):
    """Packed int32 CUDA tensors of the workload: inputs resident in HBM before the timed region starts."""
    cat = lambda arrs, w=None: torch.from_numpy(np.concatenate([a.reshape(-1) if w is None else a.reshape(-1, w) for a in arrs])).to(dev)
    return dict(ids=cat(wl["prompts"]), plen=[len(p) for p in wl["prompts"]], codes=cat(wl["spk"], 8),
                slen=[len(s) for s in wl["spk"]], text=cat(wl["text"]), tlen=[len(t) for t in wl["text"]])


PHASE_S = {"ar": 0.0, "nar": 0.0, "voc": 0.0}


def run_step_device(eng, icfg, wl, dw, T, mode, seed=0):
    """Same step with every data buffer resident on the device (mem = M5_MEM_DEVICE); the glue is slicing on device."""
    B, N, Pf, fci = wl["B"], wl["N"], wl["Pf"], wl["first_codec_idx"]
    t0 = time.perf_counter()
    eos = eng.dims["ar_vocab"] - 1
    acfg = eng.make_ar_cfg(icfg, wl["max_len"], eos, force_len=N, sync_every=64)
    out_ids, out_len, _, _ = eng.ar_generate_packed(dw["ids"], dw["plen"], dw["codes"], dw["slen"], wl["n_phones"], acfg,
                                                    seed=seed, utt=wl["utt"])
    L = int(out_len[0])  # forced length: identical for every row
    t1 = time.perf_counter()
    l0 = ((out_ids[:, fci:L] - wl["n_text"]).clamp_(min=0) % 1024).to(torch.int32).reshape(-1).contiguous()
    xlen = [L - fci] * B
    ncfg = eng.make_nar_cfg(icfg, T=T, precise=mode)
    codes = eng.nar_infer_packed(dw["text"], dw["tlen"], dw["codes"], dw["slen"], l0, xlen, ncfg, seed=seed, utt=wl["utt"])
    t2 = time.perf_counter()
    outs = codes.view(B, L - fci, 8)[:, Pf:].contiguous().view(-1, 8)
    wav = eng.vocode_packed(outs, [L - fci - Pf] * B, bandwidth_id=1)
    t3 = time.perf_counter()  # every C-ABI call returns after its stream has drained: host timers bracket device work
    PHASE_S["ar"] += t1 - t0; PHASE_S["nar"] += t2 - t1; PHASE_S["voc"] += t3 - t2
    return wav


def run_step_nar_only(eng, icfg, wl, dw, T, mode, seed=0):
    """BASELINE configs[4]: the multinomial-DDPM loop alone (L0 codes given), then the vocoder."""
    B, Pf = wl["B"], wl["Pf"]
    ncfg = eng.make_nar_cfg(icfg, T=T, precise=mode)
    xl = Pf - 1 + wl["N"]                                    # what the AR stage would hand over (prompt frames re-decoded)
    codes = eng.nar_infer_packed(dw["text"], dw["tlen"], dw["codes"], dw["slen"], dw["l0"], [xl] * B, ncfg, seed=seed, utt=wl["utt"])
    outs = codes.view(B, xl, 8)[:, Pf:].contiguous().view(-1, 8)
    return eng.vocode_packed(outs, [xl - Pf] * B, bandwidth_id=1)


# ------------------------------------------------------------------------------------------------ helpers
class ClockSampler:
    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "250"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1370.8), d.get("hbm_gbs", 6569.6), "measured (MEASURED_PEAKS.json: sustained bf16, HBM copy)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


def _thread_candidates():
    n = os.cpu_count() or 1
    return sorted({c for c in (8, 16, 32, 64, n) if c <= n})


def _best_threads(run, candidates):
    """Smallest wall time of `run()` over the candidate intra-op thread counts (a shared many-core host is often
    slowest with every logical core in the pool).  Stops growing the pool once it got 1.3x slower than the best."""
    best_t, best_n = float("inf"), candidates[0]
    for n in candidates:
        torch.set_num_threads(n)
        run()                              # first call at a new pool size pays the thread start-up
        t0 = time.perf_counter()
        run()
        t = time.perf_counter() - t0
        if t < best_t:
            best_t, best_n = t, n
        elif t > 1.3 * best_t:
            break
    return best_n, best_t


# ------------------------------------------------------------------------------------------------ CPU legs
_CPU_STATE = {}


def cpu_port_sample(size, wl, T, budget_s=25.0):
    """Times the CPU port of the reference (oracle/, fp32) on ONE utterance of the workload, bounded to ~budget_s:
    KV-cached AR decode steps at the MEAN context of the utterance (cache pre-filled with random K/V; each step includes
    the per-step speaker-encoder pass of the reference, model.py:109-127), ONE NAR model evaluation at full length
    (a reverse step = cond + uncond evaluation) and ONE posterior step (CFG + q_posterior + both Gumbel draws).  The
    prefill is charged at the measured per-token GEMM rate of the NAR evaluation.  Extrapolates
    audio_s/s = audio / (N * t_ar + T * (2 * t_fwd + t_post)).  Returns (value, detail dict)."""
    from mars5_tts_b200 import synth, weights
    from oracle import ar_oracle, nar_oracle
    t_begin = time.perf_counter()
    cands = _thread_candidates()
    key = id(size)
    if key not in _CPU_STATE:   # the synthetic checkpoints are built once per process (20 s at full size), not per sample
        _CPU_STATE[key] = (synth.make_ar_state(size), synth.make_nar_state(size))
    ar_sd, nar_sd = _CPU_STATE[key]
    cfg = weights.dims_from_state(ar_sd, nar_sd, None, size["n_text"])
    prompt, spk, text = torch.from_numpy(wl["prompts"][0]).long(), torch.from_numpy(wl["spk"][0]).long(), torch.from_numpy(wl["text"][0]).long()
    N, Pf = wl["N_b"][0], wl["Pf"]
    t_setup = time.perf_counter() - t_begin
    with torch.inference_mode():
        a, b = torch.randn(2048, 1024), torch.randn(1024, 4096)
        n_thr, _ = _best_threads(lambda: [a @ b for _ in range(4)], cands)
        torch.set_num_threads(n_thr)
        g = torch.Generator().manual_seed(0)
        # ---- AR: cached steps at the mean context
        L_mean = len(prompt) + 1 + N // 2
        H = cfg["ar_heads"]
        cache = ar_oracle.KVCache()
        for l in range(cfg["ar_layers"]):
            cache.k[l] = torch.randn(L_mean, H, 64, generator=g)
            cache.v[l] = torch.randn(L_mean, H, 64, generator=g)
        cache.n = L_mean
        t_ar_list = []
        for i in range(3):
            tok = torch.randint(size["n_text"], cfg["ar_vocab"], (1,), generator=g)
            t0 = time.perf_counter()
            ar_oracle.codeclm_step(ar_sd, cfg, tok, spk, cache)
            t_ar_list.append(time.perf_counter() - t0)
            if time.perf_counter() - t_begin - t_setup > 0.3 * budget_s:
                break
        t_ar = min(t_ar_list)
        del cache
        # ---- NAR: one evaluation; full length when it fits the budget, else a shorter S scaled by the FLOP formula
        S_full = Pf + (Pf - 1 + N)
        Tc = len(text) + 1

        def nar_flops(S):
            return 16 * (2 * 17825792 * S + 4 * S * S * 1024 + 4 * S * Tc * 1024) + 8 * 2 * 1024 * 1025 * S

        S_probe = min(256, S_full)
        x = torch.randint(0, 1025, (S_probe, 8), generator=g)
        nar_oracle.nar_forward(nar_sd, cfg, text, spk, x, T // 2, drop_cond=False)   # warms the pool, pages the weights
        t0 = time.perf_counter()
        nar_oracle.nar_forward(nar_sd, cfg, text, spk, x, T // 2, drop_cond=False)
        t_probe = time.perf_counter() - t0
        left = budget_s - (time.perf_counter() - t_begin - t_setup)
        est_full = t_probe * nar_flops(S_full) / nar_flops(S_probe)
        S = S_full if est_full < 0.8 * left else max(S_probe, int(S_full * min(1.0, 0.8 * left / max(est_full, 1e-9))))
        if S > S_probe:
            x = torch.randint(0, 1025, (S, 8), generator=g)
            t0 = time.perf_counter()
            logits = nar_oracle.nar_forward(nar_sd, cfg, text, spk, x, T // 2, drop_cond=False)
            t_fwd_S = time.perf_counter() - t0
        else:
            S, t_fwd_S = S_probe, t_probe
            logits = nar_oracle.nar_forward(nar_sd, cfg, text, spk, x, T // 2, drop_cond=False)
        t_fwd = t_fwd_S * nar_flops(S_full) / nar_flops(S)
        # ---- posterior step (diffuser.py:359-393) on the same S rows, scaled linearly in S
        tabs = nar_oracle.diffusion_tables(T)
        K = cfg["n_classes"]
        u = torch.rand(2, S, 8, K, generator=g)
        xk, m = torch.zeros_like(x), torch.zeros_like(x).bool()
        t0 = time.perf_counter()
        nar_oracle.reverse_step(tabs, logits, logits, x, xk, m, T // 2, 3.0, 0.7, u[0], u[1], K)
        t_post = (time.perf_counter() - t0) * S_full / S
        # prefill: (P+1) tokens of AR GEMM work at the FLOP rate the NAR evaluation just achieved
        rate = nar_flops(S) / t_fwd_S
        t_prefill = 2 * 687128064 * (len(prompt) + 1) / rate
    audio = (N - 1) / 75.0
    total = t_prefill + N * t_ar + T * (2 * t_fwd + t_post)
    detail = {"t_ar_step_s": round(t_ar, 4), "ar_context": L_mean, "t_nar_forward_s": round(t_fwd, 3), "nar_rows_timed": S,
              "nar_rows_full": S_full, "t_posterior_s": round(t_post, 3), "t_prefill_s_est": round(t_prefill, 3),
              "extrapolated_s_per_utterance": round(total, 1), "cores": n_thr, "host_logical_cores": os.cpu_count(),
              "sample_wall_s": round(time.perf_counter() - t_begin, 1)}
    sample = (f"1 utterance of the workload: {len(t_ar_list)} KV-cached AR steps at the mean context ({L_mean} tokens, incl. the "
              f"per-step speaker pass), 1 NAR evaluation at S={S} of {S_full} rows (x2 for cond+uncond, scaled by the FLOP formula "
              f"when S < full), 1 posterior step; extrapolated to N={N} AR steps and T={T} reverse steps")
    return audio / total, detail, sample


def gpu_eager_sample(size, wl, T, device):
    """Same-box GPU baseline (last bullet): the reference's arithmetic as plain PyTorch eager on
    THIS GPU -- the oracle port moved to the device (cuBLAS / torch kernels, none of this repo's), batch 1 like the
    reference: KV-cached AR steps at the mean context under fp16 autocast (inference.py:263), fp32 NAR evaluations
    (diffuser.py:358) and the posterior; 32 utterances = 32 sequential calls.  Bounded sample, extrapolated like
    cpu_port_sample.  A stated baseline beside the headline, not part of the product path."""
    from mars5_tts_b200 import synth, weights
    from oracle import ar_oracle, nar_oracle
    dev = torch.device("cuda", device)
    ar_sd = {k: v.to(dev) for k, v in synth.make_ar_state(size).items()}
    nar_sd = {k: v.to(dev) for k, v in synth.make_nar_state(size).items()}
    cfg = weights.dims_from_state(ar_sd, nar_sd, None, size["n_text"])
    prompt, spk, text = (torch.from_numpy(wl[k][0]).long().to(dev) for k in ("prompts", "spk", "text"))
    N, Pf = wl["N_b"][0], wl["Pf"]
    sync = torch.cuda.synchronize
    with torch.inference_mode(), torch.device(dev):
        L_mean, H = len(prompt) + 1 + N // 2, cfg["ar_heads"]
        cache = ar_oracle.KVCache()
        for l in range(cfg["ar_layers"]):
            cache.k[l] = torch.randn(L_mean, H, 64, device=dev)
            cache.v[l] = torch.randn(L_mean, H, 64, device=dev)
        cache.n = L_mean
        tok = torch.randint(size["n_text"], cfg["ar_vocab"], (1,), device=dev)
        with torch.autocast("cuda", dtype=torch.float16):
            for _ in range(2):
                ar_oracle.codeclm_step(ar_sd, cfg, tok, spk, cache)
            sync(); t0 = time.perf_counter()
            for _ in range(5):
                ar_oracle.codeclm_step(ar_sd, cfg, tok, spk, cache)
            sync(); t_ar = (time.perf_counter() - t0) / 5
        del cache
        S = Pf + (Pf - 1 + N)
        x = torch.randint(0, 1025, (S, 8), device=dev)
        nar_oracle.nar_forward(nar_sd, cfg, text, spk, x, T // 2)
        sync(); t0 = time.perf_counter()
        logits = nar_oracle.nar_forward(nar_sd, cfg, text, spk, x, T // 2)
        sync(); t_fwd = time.perf_counter() - t0
        tabs = tuple(t_.to(dev) for t_ in nar_oracle.diffusion_tables(T))
        K = cfg["n_classes"]
        u = torch.rand(2, S, 8, K, device=dev)
        xk, m = torch.zeros_like(x), torch.zeros_like(x).bool()
        sync(); t0 = time.perf_counter()
        nar_oracle.reverse_step(tabs, logits, logits, x, xk, m, T // 2, 3.0, 0.7, u[0], u[1], K)
        sync(); t_post = time.perf_counter() - t0
    total = N * t_ar + T * (2 * t_fwd + t_post)
    return ((N - 1) / 75.0) / total, {"t_ar_step_s": round(t_ar, 5), "t_nar_forward_s": round(t_fwd, 4), "t_posterior_s": round(t_post, 4),
                                     "extrapolated_s_per_utterance": round(total, 2)}


def reference_c1_sample(T_sample=2, budget_s=120.0):
    """The UNMODIFIED reference (/root/reference/mars5: ar_generate + perform_simple_inference) on BASELINE configs[0]
    (shallow clone, 6 s reference, 10-word prompt, batch 1, CPU fp32), full-size random weights; AR for a bounded number
    of tokens, NAR for T_sample of T = 200 reverse steps, extrapolated linearly (labelled).  Build container only."""
    sys.path.insert(0, "/root/reference")
    from mars5 import ar_generate as ref_ar
    from mars5 import diffuser as ref_diff
    from mars5.model import CodecLM, ResidualTransformer
    from mars5_tts_b200 import synth
    size = synth.FULL
    torch.set_num_threads(os.cpu_count() or 1)
    ar_sd, nar_sd = synth.make_ar_state(size), synth.make_nar_state(size)
    V = size["n_text"] + size["n_speech"]
    lm = CodecLM(n_vocab=V, dim=size["ar_dim"], nhead=size["ar_dim"] // 64, n_layers=size["ar_layers"],
                 n_spk_layers=size["ar_spk_layers"], dim_ff_scale=7 / 3).eval()
    nar = ResidualTransformer(n_text_vocab=size["n_text"] + 1, n_quant=1025, dim=size["nar_dim"], nhead=size["nar_dim"] // 64,
                              enc_layers=size["nar_enc_layers"], dec_layers=size["nar_dec_layers"],
                              n_spk_layers=size["nar_spk_layers"], t_emb_dim=size["nar_dim"], p_cond_drop=0, dropout=0).eval()
    lm.load_state_dict(ar_sd, strict=True); nar.load_state_dict(nar_sd, strict=True)

    class Tok:  # stand-in with the reference tokenisers' interface (the real models live in the checkpoints)
        def __init__(self, n, special):
            self.vocab, self.special_tokens = {i: (i,) for i in range(n)}, special
    tt, st = Tok(size["n_text"], {}), Tok(size["n_speech"], {"<|endofspeech|>": size["n_speech"] - 1})
    g = torch.Generator().manual_seed(0)
    Pf, n_tok, n_gen = 450, 14, 24
    spk = torch.randint(0, 1024, (Pf, 8), generator=g)
    prompt = torch.randint(0, size["n_text"] - 2, (n_tok + 2,), generator=g)
    with torch.inference_mode():
        t0 = time.perf_counter()
        seq = ref_ar.ar_generate(tt, st, lm, prompt, spk, len(prompt) + 1, max_len=len(prompt) + n_gen, fp16=False, temperature=0.7,
                                 topk=200, top_p=0.2, alpha_frequency=3, alpha_presence=0.4, penalty_window=80, eos_penalty_decay=0.5,
                                 eos_penalty_factor=1, n_phones_gen=50, vocode=False, use_kv_cache=True)
        t_ar = (time.perf_counter() - t0) / max(1, len(seq) - len(prompt))
        N = 15 * n_tok
        diff = ref_diff.MultinomialDiffusion(1025, timesteps=T_sample)
        dsh = ref_diff.DSH(last_greedy=True, x_0_temp=0.7, guidance_w=3, deep_clone=False, jump_len=1, jump_n_sample=1,
                           q0_override_steps=20, enable_kevin_scaled_inference=True, progress=False)
        _x = torch.randint(0, 1024, (1, N, 1), generator=g).repeat(1, 1, 8)
        batch = (prompt[None], spk[None].clone(), torch.tensor([len(prompt)]), torch.tensor([Pf]), _x, torch.zeros(1, N, dtype=torch.bool))
        t0 = time.perf_counter()
        ref_diff.perform_simple_inference(nar, batch, diff, T_sample, torch.float16, dsh=dsh, retain_quant0=True)
        t_nar = (time.perf_counter() - t0) / T_sample
    total = N * t_ar + 200 * t_nar
    return (N / 75.0) / total, {"t_ar_step_s": round(t_ar, 4), "t_nar_step_s": round(t_nar, 3), "cores": os.cpu_count(),
                                "extrapolated_s_per_utterance": round(total, 1)}, \
        (f"unmodified /root/reference modules, BASELINE configs[0] (shallow, Pf=450, 14 tokens -> N={N}, B=1): {n_gen - 2} AR tokens and "
         f"{T_sample} of 200 reverse steps timed, extrapolated linearly")


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", default="full", choices=["full", "mid", "tiny"])
    ap.add_argument("--config", default="c3", choices=["c2", "c3", "c4", "c5"],
                    help="BASELINE configs: c2 = deep B=1 50 tok; c3 = deep B=32 100 tok (headline); c4 = mixed lengths; c5 = NAR-only sweep")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--T", type=int, default=0)
    ap.add_argument("--mode", default=os.environ.get("M5_BENCH_MODE", ""), choices=[""] + list(MODES),
                    help="NAR numerics: fast = fp16 operands everywhere (fails the 1e-3 bound); mixed = (hi, lo) fp16 pairs for the GEMM "
                         "activations, K and V; mixed8 = mixed with the lo pass of the big decoder GEMMs in fp8; mixed8k = mixed8 with "
                         "single-fp16 keys in the decoder self-attention (all three hold 1e-3 max-abs on the logits, "
                         "tests/test_zzz_fullsize_gpu.py); precise = everything split.  Default: capi.NUM_DEFAULT, what Mars5TTS runs")
    ap.add_argument("--precise", type=int, default=-1, help="deprecated alias: 0 = fast, 1 = precise")
    ap.add_argument("--ntok", type=int, default=0, help="profiling aid: override the target-text token count (N = 15 * ntok)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--also", default="fast", help="comma list of further NAR numerics modes: one step of each is timed and reported beside "
                                                   "the headline (default: fast; skipped when the wall budget does not allow it)")
    ap.add_argument("--no-also-fast", dest="also", action="store_const", const="")
    args = ap.parse_args()
    if args.precise >= 0:
        args.mode = "precise" if args.precise else "fast"
    if not args.mode:
        from mars5_tts_b200 import capi as _capi
        args.mode = {v: k for k, v in MODES.items()}[_capi.NUM_DEFAULT]
    mode = MODES[args.mode]

    from mars5_tts_b200 import synth
    size = {"full": synth.FULL, "mid": synth.MID, "tiny": synth.TINY}[args.size]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    small = args.size != "full"
    wl_kw = dict(n_text_tok=10, n_ref_tok=5, Pf=40, frames_per_tok=4) if small else {}
    cfg_defaults = {"c2": (1, 50, 200), "c3": (32, 100, 200), "c4": (32, 100, 200), "c5": (128, 50, 64)}[args.config]
    B = args.batch or cfg_defaults[0]
    if not small:
        wl_kw["n_text_tok"] = cfg_defaults[1]
    if args.ntok:
        wl_kw["n_text_tok"] = args.ntok
    T = args.T or cfg_defaults[2]
    if args.config == "c4":
        wl_kw["mixed"] = True
    ntk, fpt = wl_kw.get("n_text_tok", 100), wl_kw.get("frames_per_tok", 15)
    names = {"c2": "BASELINE configs[1]", "c3": "BASELINE configs[2]", "c4": "BASELINE configs[3] (per-GPU shard of 32)", "c5": "BASELINE configs[4], NAR only"}
    config = {"workload": f"{'NAR-only ' if args.config == 'c5' else ''}deep-clone B={B} Pf={wl_kw.get('Pf', 450)} "
                          f"text={'U{20..120}' if args.config == 'c4' else ntk}tok N={'15*tok' if args.config == 'c4' else ntk * fpt} T={T} CFG w=3 "
                          f"({args.size} model, {names[args.config]})",
              "global_batch": B * world, "parallelism": f"dp{world}", "l2": "working set >> 126 MB L2 (weights 2.4 GB + KV 10 GB)",
              "nar_numerics": args.mode, "precise": mode}

    # -------------------------------------------------------------------------------- reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        wl = make_workload(size, 1, 1234, **wl_kw)
        vals, detail, samp, kind = [], {}, "", "port"
        t_begin = time.perf_counter()
        # every step is one bounded sample of the workload; the per-step budget is chosen so that W + K samples end within a
        # few minutes (the reference's CPU path needs ~40 minutes for ONE utterance of this workload)
        per = max(4.0, min(25.0, 200.0 / max(1, args.steps + args.warmup)))
        n_warm = 0
        for i in range(args.warmup):
            cpu_port_sample(size, wl, T, budget_s=min(per, 6.0))
            n_warm += 1
            if time.perf_counter() - t_begin > 60:
                break
        for i in range(max(1, args.steps)):
            t0 = time.perf_counter()
            v, detail, samp = cpu_port_sample(size, wl, T, budget_s=per)
            vals.append((v, time.perf_counter() - t0))
            if time.perf_counter() - t_begin > 420:  # bounded: the whole run must end within a few minutes
                break
        v = float(np.mean([a for a, _ in vals]))
        ms = float(np.mean([b for _, b in vals]) * 1e3)
        line = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": len(vals), "steps_requested": args.steps,
                "warmup": n_warm, "warmup_requested": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference", "config": config,
                "cpu_baseline": {"value": v, "unit": UNIT, "kind": kind, "sample": samp, **detail},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "note": "each step = one bounded sample of the workload on the host cores (kind: port = oracle/, the CPU restatement of "
                        "the reference; /root/reference does not exist on the GPU box)"}
        if os.path.isdir("/root/reference/mars5") and os.environ.get("M5_BENCH_REF_C1", "1") == "1":
            try:
                v1, d1, s1 = reference_c1_sample()
                line["reference_c1"] = {"value": v1, "unit": UNIT, "kind": "reference", "sample": s1, **d1}
            except Exception as e:  # the unmodified reference is an extra, never a reason to lose the line
                line["reference_c1"] = {"unavailable": repr(e)[:200]}
        print(json.dumps(line))
        return

    # -------------------------------------------------------------------------------- our arm
    from mars5_tts_b200 import dist as m5dist
    from mars5_tts_b200.engine import Engine, InferenceConfig
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as td
        td.init_process_group("nccl", device_id=torch.device("cuda", local))
    packed = m5dist.build_or_receive_weights(size, rank, world, local, max_pos=4096)
    eng = Engine(device=local, packed=packed)
    icfg = InferenceConfig()
    wl = make_workload(size, B, 1234 + rank, utt_base=rank * B, **wl_kw)
    dev = torch.device("cuda", local)
    stream = torch.cuda.ExternalStream(eng.lib.m5_stream(eng.ctx), device=dev)
    dw = to_device(wl, dev)
    nar_only = args.config == "c5"
    if nar_only:
        g = torch.Generator().manual_seed(99 + rank)
        dw["l0"] = torch.randint(0, 1024, (B * (wl["Pf"] - 1 + wl["N"]),), generator=g, dtype=torch.int32).to(dev)
    if args.config == "c4":  # mixed lengths go through the list API (per-utterance forced lengths differ): host buffers only
        raise SystemExit("config c4 is covered by tests/test_dist_*.py (sharding) -- bench lines exist for c2, c3, c5")

    def step_dev(T_=T, mode_=mode):
        wav = (run_step_nar_only if nar_only else run_step_device)(eng, icfg, wl, dw, T_, mode_)
        if world > 1:
            m5dist.all_gather_waveforms(list(wav.view(wl["B"], -1)), local)
        return wav

    def step_host():
        wavs = run_step(eng, icfg, wl, T, mode)
        if world > 1:
            m5dist.all_gather_waveforms(wavs, local)
        return wavs

    def allmax(v):
        if world == 1:
            return v
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    def timed(n, step):
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(n):
            wavs = step()
        e1.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        return allmax(e0.elapsed_time(e1)), wavs

    also = [m for m in args.also.split(",") if m and m != args.mode]
    for m in also:
        if m not in MODES:
            raise SystemExit(f"--also: unknown mode {m}")
    # ---- warm-up: W-1 short passes (full shapes, T = 8), then one complete step that calibrates the step time
    t_full_ms = None
    for i in range(args.warmup):
        if i + 1 < args.warmup:
            step_dev(T_=min(T, 8))
        else:
            t_full_ms, _ = timed(1, step_dev)
    n_steps = args.steps
    if t_full_ms is not None:
        reserve = (0 if (args.no_e2e or nar_only) else 1.08 * t_full_ms / 1e3) + (45 if (world == 1 and not args.no_cpu_baseline) else 0) + 15
        reserve += len(also) * t_full_ms / 1e3
        fit = int((BUDGET_S - allmax(elapsed()) - reserve) // (t_full_ms / 1e3))
        n_steps = max(1, min(args.steps, fit))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng.lib.m5_profile_enable(eng.ctx, 1)
    for k in PHASE_S:
        PHASE_S[k] = 0.0
    l0 = eng.launches
    ms, wav_dev = timed(n_steps, step_dev)
    launches = eng.launches - l0
    import ctypes as C
    prof = {}
    for kind, name in ((0, "gemm_tc5"), (1, "flash_attn"), (2, "ar_decode")):
        a, b, c, n = C.c_double(), C.c_double(), C.c_double(), C.c_int64()
        eng.lib.m5_profile_read(eng.ctx, kind, C.byref(a), C.byref(b), C.byref(c), C.byref(n))
        prof[name] = dict(ms=a.value, flops=b.value, bytes=c.value, launches=n.value)
    eng.lib.m5_profile_enable(eng.ctx, 0)
    clocks = sampler.stop() if rank == 0 else {}
    phases = {k: round(v / n_steps * 1e3, 1) for k, v in PHASE_S.items()}
    e2e_steps = 1  # one end-to-end step keeps the default run within minutes
    ms_e2e, wavs = (ms / n_steps, None) if (args.no_e2e or nar_only) else timed(e2e_steps, step_host)
    other_ms = {}
    for m in also:
        if t_full_ms is not None and allmax(elapsed()) + 0.8 * t_full_ms / 1e3 + 60 < BUDGET_S:
            other_ms[m], _ = timed(1, lambda: step_dev(mode_=MODES[m]))
    audio_total = wl["audio_s"] * world * n_steps
    value = audio_total / (ms / 1e3)
    if rank != 0:
        return
    peak_tf, peak_gbs, peak_src = measured_peaks()
    gp, ap_ = prof["gemm_tc5"], prof["ar_decode"]
    ach = gp["flops"] / max(gp["ms"], 1e-9) / 1e9  # TFLOP/s
    roofline = {"kernel": "gemm_tc5_2cta_kernel / gemm_tc5_kernel (tcgen05; every NAR, AR-prefill and vocoder GEMM)", "bound": "tensor",
                "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                # DRAM bytes (read + write) of ONE launch of the class's dominant shape from an ncu --set full capture
                # (profiles/r1_gemm_tc5_2cta_ncu.txt), with the algorithmic bytes of that same shape beside it (ratio 1.01);
                # `algorithmic_bytes_per_launch` is the class average the `achieved` figure is computed over
                "traffic": 376112128 + 901768704,
                "traffic_ncu": {"shape": "M=153552 N=3072 K=1024 fp16 out", "dram_bytes": 376112128 + 901768704,
                                "algorithmic_bytes": 2 * (153552 * 1024 + 3072 * 1024 + 153552 * 3072)},
                "algorithmic_bytes_per_launch": gp["bytes"] / max(gp["launches"], 1),
                "peak_source": peak_src,
                "launches": gp["launches"], "avg_launch_ms": gp["ms"] / max(gp["launches"], 1),
                "share_of_step": gp["ms"] / ms, "flash_attn_tflops": prof["flash_attn"]["flops"] / max(prof["flash_attn"]["ms"], 1e-9) / 1e9,
                "flash_attn_share_of_step": prof["flash_attn"]["ms"] / ms}
    roofline_ar = None
    if ap_["launches"] > 0:
        gbs = ap_["bytes"] / max(ap_["ms"], 1e-9) / 1e6
        roofline_ar = {"kernel": "ar_decode_kernel (one persistent cooperative kernel per decode step: all 26 layers, vocabulary projection, categorical sampler)", "bound": "hbm",
                       "achieved": gbs, "peak": peak_gbs, "unit": "GB/s", "frac": gbs / peak_gbs,
                       "bytes_per_step": ap_["bytes"] / ap_["launches"], "ms_per_decode_step": ap_["ms"] / ap_["launches"],
                       "decode_steps": ap_["launches"], "share_of_step": ap_["ms"] / ms,
                       "formula": "W_ar + sum_b 159744*(L_b+1) + B*1536*2 per step (SURVEY 8(d))",
                       # one ncu --set full capture of the kernel (profiles/r2_ar_decode_final_ncu.txt, B=32, L~1232): DRAM
                       # read + write per launch vs the algorithmic bytes of that step
                       "traffic": 7768864000 + 99603456, "traffic_algorithmic_same_launch": 7470000000}
    n_in = sum(p.nbytes for p in wl["prompts"]) + 2 * sum(s.nbytes for s in wl["spk"]) + sum(t.nbytes for t in wl["text"])
    n_out = wav_dev.numel() * 4
    e2e = {"value": wl["audio_s"] * world * e2e_steps / (ms_e2e / 1e3), "unit": UNIT, "steps": e2e_steps, "h2d_bytes_per_step": int(n_in + wl["B"] * wl["N"] * 4),
           "d2h_bytes_per_step": int(n_out + wl["B"] * wl["max_len"] * 4 + wl["B"] * (wl["Pf"] + wl["N"]) * 32),
           "note": "Engine.ar_generate / nar_infer / vocode with HOST buffers (mem=M5_MEM_HOST): ids and codes are copied in, "
                   "AR ids, NAR codes and the fp32 waveforms are copied out every step"}
    if args.no_e2e or nar_only:
        e2e["note"] = "not measured in this run (device-resident value repeated)"
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": n_steps, "steps_requested": args.steps,
           "warmup": args.warmup, "warmup_note": f"{max(args.warmup - 1, 0)} passes at T={min(T, 8)} + 1 complete step",
           "ms_per_step": ms / n_steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f16 operands (split hi/lo pairs where nar_numerics says so) / f32 accumulate", "data": "synthetic", "config": config,
           "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "roofline_ar": roofline_ar,
           "phase_ms_per_step": phases, "realtime_factor_per_gpu": value / world, "wall_s": round(elapsed(), 1)}
    if "fast" in other_ms:
        out["fast_mode"] = {"value": wl["audio_s"] * world / (other_ms["fast"] / 1e3), "unit": UNIT, "steps": 1,
                            "note": "nar_numerics=fast (fp16 operands everywhere): does NOT meet the 1e-3 max-abs logit bound"}
    for m, v in other_ms.items():
        if m != "fast":
            out.setdefault("other_modes", {})[m] = {"value": wl["audio_s"] * world / (v / 1e3), "unit": UNIT, "steps": 1, "ms_per_step": v}
    if world == 1 and not args.no_cpu_baseline:
        wl1 = make_workload(size, 1, 1234, **wl_kw)
        v, detail, samp = cpu_port_sample(size, wl1, T, budget_s=25.0)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "kind": "port", "sample": samp, **detail}
        if BUDGET_S - elapsed() > 60:
            try:   # never a reason to lose the line
                eng.close()
                del eng
                torch.cuda.empty_cache()
                gv, gd = gpu_eager_sample(size, wl1, T, local)
                out["gpu_eager_baseline"] = {"value": gv, "unit": UNIT, "kind": "oracle port in PyTorch eager on this GPU, batch 1 (fp16-autocast AR, "
                                             "fp32 NAR): the reference's own execution model, 32 utterances = 32 sequential calls", **gd}
            except Exception as e:
                out["gpu_eager_baseline"] = {"unavailable": repr(e)[:300]}
    out["wall_s"] = round(elapsed(), 1)
    print(json.dumps(out))


if __name__ == "__main__":
    try:
        main()
    finally:
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            torch.distributed.destroy_process_group()
