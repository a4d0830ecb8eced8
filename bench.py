#!/usr/bin/env python
"""bench.py -- synthesized audio seconds per second (deep-clone, batch 32) for the MARS5 AR -> NAR -> vocoder hot path.

One "step" = one pass of the hot path over one batch of B synthetic utterances (BASELINE.json configs[2]):
6 s reference clip (450 Encodec frames), 35-token reference transcript, 100-token target text, generation length forced
to N = 15 frames per text token = 1500 frames (random weights never emit EOS), T = 200 reverse steps with classifier-free
guidance.  value  = audio seconds / second with inputs resident in HBM; e2e = the same through the host-buffer C ABI.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size full|mid|tiny] [--batch B]

N > 1: launched under torch.distributed.run, one rank per GPU; utterances shard by batch (weak scaling), weights are
repacked on rank 0 and broadcast over NCCL, finished waveforms are all-gathered.  --impl reference times the CPU port of
the reference (oracle/) on the host cores on a bounded sample of the same workload (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "audio sec/s synthesized (deep-clone, batch 32)"
UNIT = "audio_s/s"


# ------------------------------------------------------------------------------------------------ workload
def make_workload(size, B, seed, n_text_tok=100, n_ref_tok=35, Pf=450, frames_per_tok=15, utt_base=0):
    """Synthetic inputs of SURVEY.md 8(d): token ids stand in for tokenised text, codes for the Encodec-encoded clip."""
    g = torch.Generator().manual_seed(seed)
    n_text = size["n_text"]
    wl = dict(B=B, Pf=Pf, N=frames_per_tok * n_text_tok, n_text=n_text, prompts=[], spk=[], text=[], n_phones=[], utt=[])
    for b in range(B):
        sot, eot = n_text - 2, n_text - 1
        text_full = [sot] + torch.randint(0, n_text - 2, (n_ref_tok + n_text_tok,), generator=g).tolist() + [eot]
        spk = torch.randint(0, 1024, (Pf, 8), generator=g).numpy().astype(np.int32)
        speech_prompt = (spk[:, 0] + n_text).tolist()  # ratio-1 synthetic speech tokenizer: one token per frame
        wl["prompts"].append(np.asarray(text_full + speech_prompt, dtype=np.int32))
        wl["spk"].append(spk)
        wl["text"].append(np.asarray(text_full, dtype=np.int32))
        wl["n_phones"].append(5 * n_text_tok)
        wl["utt"].append(utt_base + b)
    wl["first_codec_idx"] = len(wl["text"][0]) + 1          # inference.py:256
    wl["max_len"] = len(wl["prompts"][0]) + wl["N"] + 2      # generate_max_len_override so the sequence fits
    wl["audio_s"] = B * (wl["N"] - 1) / 75.0                 # frames vocoded per utterance after both crops
    return wl


def run_step(eng, icfg, wl, T, precise, seed=0):
    """AR -> (host glue of inference.py:272-283) -> NAR -> crop -> vocoder through the host-buffer API."""
    eos = eng.dims["ar_vocab"] - 1
    acfg = eng.make_ar_cfg(icfg, wl["max_len"], eos, force_len=wl["N"], sync_every=64)
    ids, _, _ = eng.ar_generate(wl["prompts"], wl["spk"], wl["n_phones"], acfg, seed=seed, utt_ids=wl["utt"])
    l0 = [((np.clip(s.astype(np.int64) - wl["n_text"], 0, None))[wl["first_codec_idx"]:] % 1024).astype(np.int32) for s in ids]
    ncfg = eng.make_nar_cfg(icfg, T=T, precise=precise)
    codes = eng.nar_infer(wl["text"], wl["spk"], l0, ncfg, seed=seed, utt_ids=wl["utt"])
    outs = [c[wl["Pf"]:] for c in codes]                     # second crop, inference.py:300-301
    return eng.vocode(outs, bandwidth_id=1)


def to_device(wl, dev):
    """Packed int32 CUDA tensors of the workload: inputs resident in HBM before the timed region starts."""
    cat = lambda arrs, w=None: torch.from_numpy(np.concatenate([a.reshape(-1) if w is None else a.reshape(-1, w) for a in arrs])).to(dev)
    return dict(ids=cat(wl["prompts"]), plen=[len(p) for p in wl["prompts"]], codes=cat(wl["spk"], 8),
                slen=[len(s) for s in wl["spk"]], text=cat(wl["text"]), tlen=[len(t) for t in wl["text"]])


PHASE_S = {"ar": 0.0, "nar": 0.0, "voc": 0.0}


def run_step_device(eng, icfg, wl, dw, T, precise, seed=0):
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
    ncfg = eng.make_nar_cfg(icfg, T=T, precise=precise)
    codes = eng.nar_infer_packed(dw["text"], dw["tlen"], dw["codes"], dw["slen"], l0, xlen, ncfg, seed=seed, utt=wl["utt"])
    t2 = time.perf_counter()
    outs = codes.view(B, L - fci, 8)[:, Pf:].contiguous().view(-1, 8)
    wav = eng.vocode_packed(outs, [L - fci - Pf] * B, bandwidth_id=1)
    t3 = time.perf_counter()  # every C-ABI call returns after its stream has drained: host timers bracket device work
    PHASE_S["ar"] += t1 - t0; PHASE_S["nar"] += t2 - t1; PHASE_S["voc"] += t3 - t2
    return wav


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
        return d.get("bf16_tflops_sustained", 1370.8), d.get("hbm_gbs", 6569.6), "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


def _thread_candidates():
    n = os.cpu_count() or 1
    return sorted({c for c in (4, 8, 16, 32, 64, n) if c <= n})


def _best_threads(run, candidates):
    """Smallest wall time of `run()` over the candidate intra-op thread counts (a shared many-core host is often
    slowest with every logical core in the pool).  Stops growing the pool once it got 1.3x slower than the best."""
    best_t, best_n = float("inf"), candidates[0]
    for n in candidates:
        torch.set_num_threads(n)
        t0 = time.perf_counter()
        run()
        t = time.perf_counter() - t0
        if t < best_t:
            best_t, best_n = t, n
        elif t > 1.3 * best_t:
            break
    return best_n, best_t


def cpu_port_sample(size, wl, T, n_ar_steps=2):
    """Times the CPU port of the reference (oracle/, fp32) on ONE utterance of the workload: prefill, a few KV-cached AR
    decode steps right after the prompt and ONE NAR forward at full length (a reverse step = cond + uncond forward).
    The intra-op thread count is chosen per phase by measurement (best of 4..all host threads).
    Extrapolates audio_s/s = audio / (prefill + N * t_ar + T * t_nar).  Returns (value, detail dict)."""
    from mars5_tts_b200 import synth, weights
    from oracle import ar_oracle, nar_oracle
    cands = _thread_candidates()
    ar_sd, nar_sd = synth.make_ar_state(size), synth.make_nar_state(size)
    cfg = weights.dims_from_state(ar_sd, nar_sd, None, size["n_text"])
    prompt, spk, text = torch.from_numpy(wl["prompts"][0]).long(), torch.from_numpy(wl["spk"][0]).long(), torch.from_numpy(wl["text"][0]).long()
    N, Pf = wl["N"], wl["Pf"]
    with torch.inference_mode():
        a, b = torch.randn(2048, 1024), torch.randn(1024, 4096)
        n_big, _ = _best_threads(lambda: [a @ b for _ in range(4)], cands)   # thread count for the GEMM-shaped phases
        torch.set_num_threads(n_big)
        t0 = time.perf_counter()
        cache = ar_oracle.KVCache()
        ar_oracle.codeclm_step(ar_sd, cfg, prompt, spk, cache)          # prefill (timed once)
        t_prefill = time.perf_counter() - t0
        g = torch.Generator().manual_seed(0)

        def ar_steps():                                                  # KV-cached steps right after the prompt (the
            for _ in range(n_ar_steps):                                  # cheapest context of the utterance), incl. the
                tok = torch.randint(size["n_text"], cfg["ar_vocab"], (1,), generator=g)  # per-step speaker pass of the
                ar_oracle.codeclm_step(ar_sd, cfg, tok, spk, cache)      # reference
        n_small, t_steps = _best_threads(ar_steps, cands)
        t_ar = t_steps / n_ar_steps
        torch.set_num_threads(n_big)
        S = Pf + (Pf - 1 + N)
        x = torch.randint(0, 1025, (S, 8), generator=g)
        t0 = time.perf_counter()
        nar_oracle.nar_forward(nar_sd, cfg, text, spk, x, T // 2, drop_cond=False)
        t_nar = 2.0 * (time.perf_counter() - t0)                         # a reverse step = cond + uncond forward (+ posterior, not counted)
    audio = (N - 1) / 75.0
    total = t_prefill + N * t_ar + T * t_nar
    detail = {"t_prefill_s": round(t_prefill, 3), "t_ar_step_s": round(t_ar, 4), "t_nar_step_s": round(t_nar, 3),
              "extrapolated_s_per_utterance": round(total, 1), "threads_gemm_phases": n_big, "threads_ar_steps": n_small,
              "cores": max(n_big, n_small)}
    return audio / total, detail


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", default="full", choices=["full", "mid", "tiny"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--T", type=int, default=200)
    ap.add_argument("--precise", type=int, default=0)
    ap.add_argument("--ntok", type=int, default=0, help="profiling aid: override the target-text token count (N = 15 * ntok)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    from mars5_tts_b200 import synth
    size = {"full": synth.FULL, "mid": synth.MID, "tiny": synth.TINY}[args.size]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    small = args.size != "full"
    wl_kw = dict(n_text_tok=10, n_ref_tok=5, Pf=40, frames_per_tok=4) if small else {}
    if args.ntok:
        wl_kw["n_text_tok"] = args.ntok
    config = {"workload": f"deep-clone B={args.batch} Pf={wl_kw.get('Pf', 450)} text={wl_kw.get('n_text_tok', 100)}tok "
                          f"N={wl_kw.get('n_text_tok', 100) * wl_kw.get('frames_per_tok', 15)} T={args.T} CFG w=3 ({args.size} model, "
                          f"BASELINE configs[2])",
              "global_batch": args.batch * world, "parallelism": f"dp{world}", "l2": "working set >> 126 MB L2 (weights 2.4 GB + KV 10 GB)",
              "precise": args.precise}

    # -------------------------------------------------------------------------------- reference arm (CPU port)
    if args.impl == "reference":
        if rank != 0:
            return
        wl = make_workload(size, 1, 1234, **wl_kw)
        vals, detail = [], {}
        n_warm = 0  # deterministic CPU work (no clocks to settle): warm-up passes would only push the run past minutes
        t_begin = time.perf_counter()
        for i in range(n_warm + args.steps):
            t0 = time.perf_counter()
            v, detail = cpu_port_sample(size, wl, args.T, n_ar_steps=2)
            if i >= n_warm:
                vals.append((v, time.perf_counter() - t0))
            if vals and time.perf_counter() - t_begin > 150:  # bounded: the whole run must end within a few minutes
                break
        config["steps_measured"] = len(vals)
        v = float(np.mean([a for a, _ in vals])) if vals else 0.0
        ms = float(np.mean([b for _, b in vals]) * 1e3) if vals else 0.0
        samp = "1 utterance: prefill + 2 KV-cached AR steps (context = prompt) + 1 NAR forward at S=2399 (x2 for cond+uncond), extrapolated to N AR steps and T reverse steps"
        print(json.dumps({"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                          "data": "synthetic", "impl": "reference", "config": config,
                          "cpu_baseline": {"value": v, "unit": UNIT, "kind": "port", "sample": samp, **detail},
                          "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
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
    wl = make_workload(size, args.batch, 1234 + rank, utt_base=rank * args.batch, **wl_kw)
    stream = torch.cuda.ExternalStream(eng.lib.m5_stream(eng.ctx), device=torch.device("cuda", local))

    dw = to_device(wl, torch.device("cuda", local))

    def step_dev():
        wav = run_step_device(eng, icfg, wl, dw, args.T, args.precise)
        if world > 1:
            m5dist.all_gather_waveforms(list(wav.view(wl["B"], -1)), local)
        return wav

    def step_host():
        wavs = run_step(eng, icfg, wl, args.T, args.precise)
        if world > 1:
            m5dist.all_gather_waveforms(wavs, local)
        return wavs

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
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=f"cuda:{local}")
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t.item())
        return ms, wavs

    for _ in range(args.warmup):
        step_dev()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng.lib.m5_profile_enable(eng.ctx, 1)
    for k in PHASE_S:
        PHASE_S[k] = 0.0
    l0 = eng.launches
    ms, wav_dev = timed(args.steps, step_dev)
    launches = eng.launches - l0
    import ctypes as C
    prof = {}
    for kind, name in ((0, "gemm_tc5"), (1, "flash_attn")):
        a, b, c, n = C.c_double(), C.c_double(), C.c_double(), C.c_int64()
        eng.lib.m5_profile_read(eng.ctx, kind, C.byref(a), C.byref(b), C.byref(c), C.byref(n))
        prof[name] = dict(ms=a.value, flops=b.value, bytes=c.value, launches=n.value)
    eng.lib.m5_profile_enable(eng.ctx, 0)
    clocks = sampler.stop() if rank == 0 else {}
    phases = {k: round(v / args.steps * 1e3, 1) for k, v in PHASE_S.items()}
    e2e_steps = 1  # one end-to-end step keeps the default run within minutes (each step is ~40 s of GPU work)
    ms_e2e, wavs = (ms / args.steps, None) if args.no_e2e else timed(e2e_steps, step_host)
    audio_total = wl["audio_s"] * world * args.steps
    value = audio_total / (ms / 1e3)
    if rank != 0:
        return
    peak_tf, peak_gbs, peak_src = measured_peaks()
    gp = prof["gemm_tc5"]
    ach = gp["flops"] / max(gp["ms"], 1e-9) / 1e9  # TFLOP/s
    roofline = {"kernel": "gemm_tc5_2cta_kernel / gemm_tc5_kernel (tcgen05; every NAR, AR-prefill and vocoder GEMM)", "bound": "tensor",
                "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                # `achieved` averages over every GEMM shape of the step, so no single DRAM-byte figure matches it; the one
                # ncu --set full capture (profiles/r1_gemm_tc5_2cta_ncu.txt) is reported with its own shape instead
                "traffic": None,
                "traffic_ncu": {"shape": "M=153552 N=3072 K=1024 fp16 out", "dram_bytes": 376112128 + 901768704,
                                "algorithmic_bytes": 2 * (153552 * 1024 + 3072 * 1024 + 153552 * 3072)},
                "peak_source": peak_src,
                "launches": gp["launches"], "avg_launch_ms": gp["ms"] / max(gp["launches"], 1),
                "share_of_step": gp["ms"] / ms, "flash_attn_tflops": prof["flash_attn"]["flops"] / max(prof["flash_attn"]["ms"], 1e-9) / 1e9,
                "flash_attn_share_of_step": prof["flash_attn"]["ms"] / ms}
    n_in = sum(p.nbytes for p in wl["prompts"]) + 2 * sum(s.nbytes for s in wl["spk"]) + sum(t.nbytes for t in wl["text"])
    n_out = wav_dev.numel() * 4
    e2e = {"value": wl["audio_s"] * world * e2e_steps / (ms_e2e / 1e3), "unit": UNIT, "steps": e2e_steps, "h2d_bytes_per_step": int(n_in + wl["B"] * wl["N"] * 4),
           "d2h_bytes_per_step": int(n_out + wl["B"] * wl["max_len"] * 4 + wl["B"] * (wl["Pf"] + wl["N"]) * 32),
           "note": "Engine.ar_generate / nar_infer / vocode with HOST buffers (mem=M5_MEM_HOST): ids and codes are copied in, "
                   "AR ids, NAR codes and the fp32 waveforms are copied out every step"}
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f16 operands / f32 accumulate", "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e,
           "gpu_launches": int(launches), "roofline": roofline, "phase_ms_per_step": phases, "realtime_factor_per_gpu": value / world}
    if world == 1 and not args.no_cpu_baseline:
        v, detail = cpu_port_sample(size, make_workload(size, 1, 1234, **wl_kw), args.T, n_ar_steps=2)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "kind": "port",
                               "sample": "1 utterance: prefill + 2 KV-cached AR steps (context = prompt) + 1 NAR forward at S=2399 (x2), extrapolated to N AR steps and T reverse steps", **detail}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
