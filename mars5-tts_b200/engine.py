"""Host-side mirror of the reference interface for the hot path.

* ``Engine``      -- thin, typed wrapper over the C ABI (include/mars5_b200.h): batches of independent utterances in,
                     ids / codes / waveforms out.  All arithmetic happens in libmars5_b200.so; this file only packs
                     pointers.  There is no PyTorch/CPU fallback: a missing library or GPU raises.
* ``InferenceConfig`` / ``Mars5TTS`` -- drop-in for /root/reference/inference.py:24-77 and :80-307
                     (same constructor inputs ``ar_ckpt`` / ``nar_ckpt``, same ``tts()`` / ``vocode()`` signatures and
                     return values).  Tokenisers and the Encodec encoder are outside the hot path (SURVEY.md section 2 rows
                     10-11); they are taken from the reference package when it is importable or injected by the caller.
"""
import ctypes as C
import logging
from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch

from . import capi
from . import weights
from .trim import trim_bounds_batch


@dataclass
class InferenceConfig:
    """Same fields and defaults as the reference's InferenceConfig (inference.py:24-77)."""
    temperature: float = 0.7
    top_k: int = 200
    top_p: float = 0.2
    typical_p: float = 1.0
    freq_penalty: float = 3
    presence_penalty: float = 0.4
    rep_penalty_window: int = 80
    eos_penalty_decay: float = 0.5
    eos_penalty_factor: float = 1
    eos_estimated_gen_length_factor: float = 1.0
    timesteps: int = 200
    x_0_temp: float = 0.7
    q0_override_steps: int = 20
    nar_guidance_w: float = 3
    max_prompt_dur: float = 12
    generate_max_len_override: int = -1
    deep_clone: bool = True
    use_kv_cache: bool = True
    trim_db: float = 27
    beam_width: int = 1
    ref_audio_pad: float = 0


class _phase:
    """NVTX range around one C-ABI call (tracing hook): the phase (AR generate, NAR reverse loop, vocoder, ...)
    shows up by name in an ncu / nsys timeline.  A no-op when torch was built without NVTX."""
    _ok = True

    def __init__(self, name):
        self.name, self.pushed = name, False

    def __enter__(self):
        if _phase._ok:
            try:
                torch.cuda.nvtx.range_push(self.name)
                self.pushed = True
            except Exception:
                _phase._ok = False
        return self

    def __exit__(self, *exc):
        if self.pushed:
            try:
                torch.cuda.nvtx.range_pop()
            except Exception:
                _phase._ok = False
        return False


def _i32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


def _cat_i32(arrs, width=None):
    if len(arrs) == 0:
        return np.zeros((0,) if width is None else (0, width), dtype=np.int32)
    return np.ascontiguousarray(np.concatenate([np.asarray(a, dtype=np.int32).reshape(-1) if width is None
                                                else np.asarray(a, dtype=np.int32).reshape(-1, width) for a in arrs]))


class Engine:
    """One context on one GPU.  `tensors` (already-repacked CPU tensors) may be given to skip the repack (multi-GPU:
    rank 0 repacks, the blob is broadcast over NCCL, see dist.py)."""

    def __init__(self, ar_sd=None, nar_sd=None, voc_sd=None, text_vocab_len=None, device=0, max_pos=4096,
                 packed=None, enc_sd=None):
        if not torch.cuda.is_available():
            raise RuntimeError("mars5_tts_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = capi.load()
        self.device = int(device)
        if packed is None:
            dims = weights.dims_from_state(ar_sd, nar_sd, voc_sd, text_vocab_len)
            tensors, alphas = weights.repack(ar_sd, nar_sd, voc_sd, dims, max_pos=max_pos)
            packed = {"dims": dims, "alphas": alphas, "tensors": tensors, "max_pos": max_pos}
        if enc_sd is not None:   # Encodec encoder + RVQ weights: enables encodec_encode()
            packed = dict(packed, tensors={**packed["tensors"], **weights.repack_encodec(enc_sd)})
        self.has_encodec = any(k.startswith("enc.") for k in packed["tensors"])
        self.dims, self.alphas = packed["dims"], packed["alphas"]
        dev = torch.device("cuda", self.device)
        self.tensors = {k: (v if v.is_cuda else v.to(dev)) for k, v in packed["tensors"].items()}
        self.cfg = weights.make_cfg(self.dims, self.alphas, packed["max_pos"])
        arr = (capi.Tensor * len(self.tensors))()
        self._names = []
        for i, (k, v) in enumerate(self.tensors.items()):
            nm = k.encode()
            self._names.append(nm)
            arr[i].name, arr[i].ptr, arr[i].numel = nm, v.data_ptr(), v.numel()
            arr[i].dtype = capi.DT_F16 if v.dtype == torch.float16 else (capi.DT_U8 if v.dtype == torch.uint8 else capi.DT_F32)
        self.ctx = C.c_void_p()
        rc = self.lib.m5_create(self.device, C.byref(self.cfg), arr, len(self.tensors), C.byref(self.ctx))
        if rc != 0:
            raise capi.M5Error(f"m5_create failed with code {rc}")
        self._sched_cache = {}

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.m5_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self):
        return int(self.lib.m5_launch_count(self.ctx))

    # ------------------------------------------------------------------------------------------ helpers
    @staticmethod
    def _mem(x):
        return capi.MEM_DEVICE if torch.is_tensor(x) and x.is_cuda else capi.MEM_HOST

    def _fence(self, mem):
        """Device buffers (inputs made by the caller, outputs zero-filled below) are produced on torch's current stream but
        consumed on the context's own stream: drain torch's stream first.  Every C-ABI call returns after ITS stream has
        drained, so results are safe to read from torch afterwards."""
        if mem == capi.MEM_DEVICE:
            torch.cuda.current_stream(torch.device("cuda", self.device)).synchronize()

    def _alloc(self, mem, shape, dtype):
        if mem == capi.MEM_DEVICE:
            return torch.zeros(shape, dtype={np.int32: torch.int32, np.float32: torch.float32}[dtype],
                               device=torch.device("cuda", self.device))
        return np.zeros(shape, dtype=dtype)

    # ------------------------------------------------------------------------------------------ AR
    def make_ar_cfg(self, icfg: InferenceConfig, max_len, eos_id, force_len=0, sync_every=16):
        c = capi.ArCfg()
        c.temperature, c.top_k, c.top_p = icfg.temperature, icfg.top_k, icfg.top_p
        c.alpha_frequency, c.alpha_presence, c.penalty_window = icfg.freq_penalty, icfg.presence_penalty, icfg.rep_penalty_window
        c.eos_penalty_decay, c.eos_penalty_factor = icfg.eos_penalty_decay, icfg.eos_penalty_factor
        c.max_len, c.eos_id, c.force_len, c.sync_every = max_len, eos_id, force_len, sync_every
        c.typical_p = icfg.typical_p
        return c

    def ar_generate_packed(self, ids, plen, codes, slen, nph, ar_cfg, noise=None, seed=0, utt=None, dump_steps=0):
        """Packed form: `ids` [sum plen] and `codes` [sum slen, 8] are int32 numpy arrays (HOST) or CUDA tensors (DEVICE);
        plen / slen / nph / utt are host arrays.  Returns (out_ids [B, max_len], out_len, hit, dump) in the same memory."""
        mem, B = self._mem(ids), len(plen)
        # keep every temporary alive until the C call returns (ctypes only sees raw addresses)
        plen_a, slen_a = _i32(plen), _i32(slen)
        nph_a = _i32(nph) if nph is not None else None
        utt_a = np.ascontiguousarray(np.asarray(utt, dtype=np.int64)) if utt is not None else None
        V, max_len = self.dims["ar_vocab"], ar_cfg.max_len
        out_ids = self._alloc(mem, (B, max_len), np.int32)
        out_len, hit = np.zeros(B, dtype=np.int32), np.zeros(B, dtype=np.int32)
        nsteps = 0 if noise is None else noise.shape[1]
        dump = self._alloc(mem, (B, dump_steps, V), np.float32) if dump_steps else None
        self._fence(mem)
        with _phase("m5_ar_generate"):
            rc = self.lib.m5_ar_generate(self.ctx, B, capi.ptr(ids), capi.ptr(plen_a), capi.ptr(codes), capi.ptr(slen_a),
                                         capi.ptr(nph_a), C.byref(ar_cfg), mem, capi.ptr(noise), nsteps, C.c_uint64(seed),
                                         capi.ptr(utt_a), capi.ptr(out_ids), capi.ptr(out_len), capi.ptr(hit), capi.ptr(dump),
                                         dump_steps)
        capi.check(self.ctx, rc, "m5_ar_generate")
        return out_ids, out_len, hit, dump

    def ar_generate(self, prompts, spk_codes, n_phones, ar_cfg, noise=None, seed=0, utt_ids=None, dump_steps=0):
        """prompts: list of int sequences; spk_codes: list of (Pf, 8) arrays (host buffers).
        Returns (list of id arrays, hit_maxlen list, logits dump or None)."""
        nz = np.ascontiguousarray(np.asarray(noise, dtype=np.float32)) if noise is not None else None
        out_ids, out_len, hit, dump = self.ar_generate_packed(_cat_i32(prompts), [len(p) for p in prompts], _cat_i32(spk_codes, 8),
                                                              [len(s) for s in spk_codes], n_phones, ar_cfg, nz, seed, utt_ids,
                                                              dump_steps)
        return [out_ids[b, :out_len[b]].copy() for b in range(len(prompts))], hit.tolist(), dump

    def ar_forward(self, prompts, spk_codes):
        B = len(prompts)
        ids, plen = _cat_i32(prompts), _i32([len(p) for p in prompts])
        codes, slen = _cat_i32(spk_codes, 8), _i32([len(s) for s in spk_codes])
        V = self.dims["ar_vocab"]
        out = np.zeros((int(plen.sum()), V), dtype=np.float32)
        with _phase("m5_ar_forward"):
            rc = self.lib.m5_ar_forward(self.ctx, B, capi.ptr(ids), capi.ptr(plen), capi.ptr(codes), capi.ptr(slen),
                                        capi.MEM_HOST, capi.ptr(out))
        capi.check(self.ctx, rc, "m5_ar_forward")
        offs = np.concatenate([[0], np.cumsum(plen)])
        return [out[offs[b]:offs[b + 1]] for b in range(B)]

    # ------------------------------------------------------------------------------------------ NAR
    def schedule(self, T):
        if T not in self._sched_cache:
            self._sched_cache[T] = np.ascontiguousarray(weights.diffusion_schedule(T).numpy())
        return self._sched_cache[T]

    def make_nar_cfg(self, icfg: InferenceConfig, T=200, precise=capi.NUM_DEFAULT, jump_len=1, jump_n_sample=1, scaled_forward=False):
        """precise: NAR numerics (capi.NUM_FAST / NUM_PRECISE / NUM_MIXED / NUM_MIXED8 / NUM_MIXED8K; True == NUM_PRECISE).  The
        default (capi.NUM_DEFAULT) is a setting whose logits stay within 1e-3 max-abs of the fp32 reference at full dims
        (tests/test_zzz_fullsize_gpu.py); the fp8 modes are `mixed` wherever the fp8 pass does not apply (short sequences,
        small batches)."""
        c = capi.NarCfg()
        c.T, c.x0_temp, c.guidance_w = T, icfg.x_0_temp, icfg.nar_guidance_w
        c.q0_override_steps, c.deep_clone, c.precise = icfg.q0_override_steps, int(icfg.deep_clone), int(precise)
        c.schedule = self.schedule(T).ctypes.data
        c.jump_len, c.jump_n_sample, c.scaled_forward = int(jump_len), int(jump_n_sample), int(scaled_forward)
        return c

    def nar_infer_packed(self, text, tlen, codes, clen, l0, xlen, nar_cfg, x_init=None, noise=None, seed=0, utt=None):
        """Packed form (numpy = HOST, CUDA tensors = DEVICE).  Returns codes [sum xlen, 8] in the same memory."""
        mem, B = self._mem(text), len(tlen)
        tlen_a, clen_a, xlen_a = _i32(tlen), _i32(clen), _i32(xlen)
        utt_a = np.ascontiguousarray(np.asarray(utt, dtype=np.int64)) if utt is not None else None
        out = self._alloc(mem, (int(np.sum(xlen)), 8), np.int32)
        self._fence(mem)
        with _phase("m5_nar_infer"):
            rc = self.lib.m5_nar_infer(self.ctx, B, capi.ptr(text), capi.ptr(tlen_a), capi.ptr(codes), capi.ptr(clen_a),
                                       capi.ptr(l0), capi.ptr(xlen_a), C.byref(nar_cfg), mem, capi.ptr(x_init), capi.ptr(noise),
                                       C.c_uint64(seed), capi.ptr(utt_a), capi.ptr(out))
        capi.check(self.ctx, rc, "m5_nar_infer")
        return out

    def nar_infer(self, c_text, c_codes, x_l0, nar_cfg, x_init=None, noise=None, seed=0, utt_ids=None):
        xlen = [len(x) for x in x_l0]
        xi = _cat_i32(x_init, 8) if x_init is not None else None
        nz = np.ascontiguousarray(np.asarray(noise, dtype=np.float32)) if noise is not None else None
        out = self.nar_infer_packed(_cat_i32(c_text), [len(t) for t in c_text], _cat_i32(c_codes, 8), [len(c) for c in c_codes],
                                    _cat_i32(x_l0), xlen, nar_cfg, xi, nz, seed, utt_ids)
        offs = np.concatenate([[0], np.cumsum(xlen)])
        return [out[offs[b]:offs[b + 1]] for b in range(len(c_text))]

    def nar_forward(self, c_text, c_codes, x, t, drop_cond=False, precise=capi.NUM_DEFAULT):
        B = len(c_text)
        text, tlen = _cat_i32(c_text), _i32([len(v) for v in c_text])
        codes, clen = _cat_i32(c_codes, 8), _i32([len(v) for v in c_codes])
        xs, xlen = _cat_i32(x, 8), _i32([len(v) for v in x])
        K = self.dims["n_classes"]
        out = np.zeros((int(xlen.sum()), 8, K), dtype=np.float32)
        with _phase("m5_nar_forward"):
            rc = self.lib.m5_nar_forward(self.ctx, B, capi.ptr(text), capi.ptr(tlen), capi.ptr(codes), capi.ptr(clen),
                                         capi.ptr(xs), capi.ptr(xlen), int(t), int(drop_cond), int(precise), capi.MEM_HOST,
                                         capi.ptr(out))
        capi.check(self.ctx, rc, "m5_nar_forward")
        offs = np.concatenate([[0], np.cumsum(xlen)])
        return [out[offs[b]:offs[b + 1]] for b in range(B)]

    # ------------------------------------------------------------------------------------------ Encodec encoder
    def encodec_encode(self, wavs, n_q=8):
        """wavs: list of mono 24 kHz waveforms (1-D float arrays / tensors).  Returns a list of (T_b, n_q) int32 code arrays,
        T_b = ceil(len / 320) -- what EncodecModel.encode gives for each clip (inference.py:233), for the whole batch at once."""
        if not self.has_encodec:
            raise capi.M5Error("this Engine was built without Encodec weights (pass enc_sd=EncodecModel.state_dict())")
        arrs = [np.ascontiguousarray(np.asarray(w.detach().cpu() if torch.is_tensor(w) else w, dtype=np.float32).reshape(-1)) for w in wavs]
        ns = _i32([len(a) for a in arrs])
        flat = np.ascontiguousarray(np.concatenate(arrs))
        frames = [(int(n) + 319) // 320 for n in ns]
        out = np.zeros((int(np.sum(frames)), n_q), dtype=np.int32)
        with _phase("m5_encodec_encode"):
            rc = self.lib.m5_encodec_encode(self.ctx, len(arrs), capi.ptr(flat), capi.ptr(ns), capi.MEM_HOST, int(n_q), capi.ptr(out))
        capi.check(self.ctx, rc, "m5_encodec_encode")
        offs = np.concatenate([[0], np.cumsum(frames)])
        return [out[offs[b]:offs[b + 1]] for b in range(len(arrs))]

    # ------------------------------------------------------------------------------------------ vocoder
    def vocode_packed(self, codes, n_frames, bandwidth_id=1):
        mem, nf_a = self._mem(codes), _i32(n_frames)
        out = self._alloc(mem, (int(np.sum(n_frames)) * self.dims["voc_hop"],), np.float32)
        self._fence(mem)
        with _phase("m5_vocode"):
            rc = self.lib.m5_vocode(self.ctx, len(n_frames), capi.ptr(codes), capi.ptr(nf_a), int(bandwidth_id), mem, capi.ptr(out))
        capi.check(self.ctx, rc, "m5_vocode")
        return out

    def vocode_trim(self, codes, top_db, bandwidth_id=1, frame_length=2048, hop_length=512):
        """Vocoder + the reference's silence trim (inference.py:304-305) in one call; the frame powers are computed on the
        device behind the overlap-add.  Returns (list of untrimmed waveforms, list of (start, end) sample ranges)."""
        nf = [len(c) for c in codes]
        cat, nf_a = _cat_i32(codes, 8), _i32(nf)
        out = np.zeros((int(np.sum(nf)) * self.dims["voc_hop"],), dtype=np.float32)
        st, en = np.zeros(len(nf), dtype=np.int64), np.zeros(len(nf), dtype=np.int64)
        with _phase("m5_vocode_trim"):
            rc = self.lib.m5_vocode_trim(self.ctx, len(nf), capi.ptr(cat), capi.ptr(nf_a), int(bandwidth_id), capi.MEM_HOST, float(top_db),
                                         int(frame_length), int(hop_length), capi.ptr(out), capi.ptr(st), capi.ptr(en))
        capi.check(self.ctx, rc, "m5_vocode_trim")
        offs = np.concatenate([[0], np.cumsum(nf)]) * self.dims["voc_hop"]
        return [out[offs[b]:offs[b + 1]] for b in range(len(codes))], [(int(a), int(b)) for a, b in zip(st, en)]

    def vocode(self, codes, bandwidth_id=1):
        nf = [len(c) for c in codes]
        out = self.vocode_packed(_cat_i32(codes, 8), nf, bandwidth_id)
        offs = np.concatenate([[0], np.cumsum(nf)]) * self.dims["voc_hop"]
        return [out[offs[b]:offs[b + 1]] for b in range(len(codes))]


class Mars5TTS:
    """Drop-in for the reference's Mars5TTS (inference.py:79-307) on the hot path.

    ``ar_ckpt`` / ``nar_ckpt``: the dicts hubconf.py builds ({'vocab': {...}, 'model': state_dict}).  The text /
    speech tokenisers are built from the "minbpe v1" models in ``ar_ckpt['vocab']`` (mars5_tts_b200.bpe, native merge
    engine) or passed in (``texttok`` / ``speechtok``: anything with the reference tokenisers' interface); ``codec`` must offer ``encode(wav[None]) -> [(codes (1, 8, T), scale)]`` like
    EncodecModel; ``vocos_state`` is the state dict of Vocos("charactr/vocos-encodec-24khz") (weight-norm removed).
    """

    def __init__(self, ar_ckpt, nar_ckpt, device: Optional[str] = None, *, vocos_state=None, texttok=None,
                 speechtok=None, codec=None, encodec_state=None):
        if texttok is None or speechtok is None:
            # "minbpe v1" model texts stored in the checkpoint (inference.py:92-99), on the native merge engine (bpe.py)
            from . import bpe
            texttok = bpe.RegexTokenizer(bpe.GPT4_SPLIT_PATTERN)
            texttok.load(ar_ckpt["vocab"]["texttok.model"])
            speechtok = bpe.CodebookTokenizer(bpe.GPT4_SPLIT_PATTERN)
            speechtok.load(ar_ckpt["vocab"]["speechtok.model"])
        dev_index = torch.device(device).index if device not in (None, "cuda") else None
        self.device = torch.device("cuda", dev_index or 0)
        if codec is None:
            # like the reference (inference.py:87-88): EncodecModel.encodec_model_24khz() at 6 kbps on the device.  The Encodec
            # ENCODER is outside the hot path (third-party package); it is only needed by tts().
            # When the package imports, only its WEIGHTS are taken: the encoder + RVQ run in libmars5_b200.so (csrc/encodec.cu,
            # batched over the reference clips); `encodec_state=` passes such a state dict directly.
            if encodec_state is None:
                try:
                    from encodec import EncodecModel
                    encodec_state = EncodecModel.encodec_model_24khz().state_dict()
                except ImportError:
                    encodec_state = None   # tts() raises with instructions; vocode() / the Engine keep working
        if vocos_state is None:
            # like the reference (inference.py:119-121): Vocos.from_pretrained("charactr/vocos-encodec-24khz"), weight norm
            # folded; only its state dict is used -- the network itself runs in libmars5_b200.so
            try:
                from vocos import Vocos
                vocos_state = Vocos.from_pretrained("charactr/vocos-encodec-24khz").state_dict()
            except ImportError:
                raise RuntimeError("the `vocos` package is not importable: pass vocos_state=<state dict of "
                                   "Vocos.from_pretrained('charactr/vocos-encodec-24khz')> to Mars5TTS (the vocoder network runs "
                                   "inside libmars5_b200.so, only its weights are needed)") from None
        self.texttok, self.speechtok, self.codec = texttok, speechtok, codec
        self.engine = Engine(ar_ckpt["model"], nar_ckpt["model"], vocos_state, len(texttok.vocab), device=self.device.index,
                             enc_sd=encodec_state if codec is None else None)
        self.n_vocab = len(texttok.vocab) + len(speechtok.vocab)
        self.n_text_vocab = len(texttok.vocab) + 1
        self.diffusion_n_classes = 1025
        self.default_T = 200
        self.sr, self.latent_sr = 24000, 75
        self._calls = 0   # tts() draws fresh randomness on every call, like the reference's unseeded torch generator

    @torch.inference_mode()
    def vocode(self, tokens: torch.Tensor) -> torch.Tensor:
        """tokens (seq_len, n_q) -> (1, T) float CPU tensor, bandwidth_id = 1 like the reference (inference.py:165-171)."""
        wav = self.engine.vocode([tokens.detach().cpu().numpy()], bandwidth_id=1)[0]
        return torch.from_numpy(wav)[None]

    def get_speaker_embedding(self, ref_audio: torch.Tensor) -> torch.Tensor:
        """The reference's analysis helper (inference.py:174-199: position 0 of the AR speaker encoder's output for a clip) is NOT
        part of the hot path and has no entry point in the C ABI: the library computes that vector inside m5_ar_generate and never
        exposes it.  Raising beats returning something different from the reference."""
        raise NotImplementedError("Mars5TTS.get_speaker_embedding is outside the accelerated path (SURVEY.md section 8): compute it "
                                  "with the reference's CodecLM (inference.py:174-199); tts() / tts_batch() / vocode() do not need it")

    def _prepare(self, text, ref_audio, ref_transcript, cfg):
        """Host glue of inference.py:212-258: tokenise, encode the reference clip, build the AR prompt."""
        if cfg.deep_clone and ref_transcript is None:
            raise AssertionError("Inference config deep clone is set to true, but reference transcript not specified! "
                                 "Please specify the transcript of the prompt, or set deep_clone=False in the inference `cfg` argument.")
        if ref_audio.shape[-1] / self.sr > cfg.max_prompt_dur:
            logging.warning("Reference audio duration is > max suggested ref audio. Expect quality degradations.")
        tt = self.texttok
        text_tokens = tt.encode("<|startoftext|>" + text.strip() + "<|endoftext|>", allowed_special="all")
        if ref_audio.dim() == 1:
            ref_audio = ref_audio[None]
        if ref_audio.shape[0] != 1:
            ref_audio = ref_audio.mean(dim=0, keepdim=True)
        ref_audio = torch.nn.functional.pad(ref_audio, (int(self.sr * cfg.ref_audio_pad), 0))
        if self.codec is None and self.engine.has_encodec:
            codes = self.engine.encodec_encode([ref_audio[0]])[0]          # (T, 8) on the device-side encoder
            prompt_codec = torch.from_numpy(codes.astype(np.int64)).T[None]  # (1, n_q, T) like EncodecModel.encode
            return self._prepare_with_codes(text, ref_transcript, cfg, prompt_codec)
        if self.codec is None:
            raise RuntimeError("no Encodec codec: `encodec` is not importable and neither codec= nor encodec_state= was passed to "
                               "Mars5TTS (codec: anything with encode(wav[None]) -> [(codes (1, 8, T), scale)])")
        wav_in = ref_audio[None]
        if hasattr(self.codec, "parameters"):   # an nn.Module codec lives on its own device (the reference moves the clip there)
            p0 = next(iter(self.codec.parameters()), None)
            if p0 is not None:
                wav_in = wav_in.to(p0.device)
        prompt_codec = self.codec.encode(wav_in)[0][0]  # (1, n_q, T)
        return self._prepare_with_codes(text, ref_transcript, cfg, prompt_codec)

    def _prepare_with_codes(self, text, ref_transcript, cfg, prompt_codec):
        tt = self.texttok
        text_tokens = tt.encode("<|startoftext|>" + text.strip() + "<|endoftext|>", allowed_special="all")
        l0 = prompt_codec[0, 0].tolist()
        speech_tokens = self.speechtok.encode(" ".join(str(t) for t in l0).strip())
        spk_ref = prompt_codec[0].T.cpu().numpy().astype(np.int32)  # (T, n_q)
        offset_codes = [p + len(tt.vocab) for p in speech_tokens]
        n_speech_inp = 0
        if not cfg.deep_clone:
            offset_codes = offset_codes[:0]
        else:
            text_tokens = tt.encode("<|startoftext|>" + ref_transcript + " " + str(text).strip() + "<|endoftext|>",
                                    allowed_special="all")
            n_speech_inp = len(offset_codes)
        prompt = text_tokens + offset_codes
        first_codec_idx = len(prompt) - n_speech_inp + 1
        return dict(prompt=prompt, first_codec_idx=first_codec_idx, text_tokens=text_tokens, spk_ref=spk_ref,
                    n_phones=round(cfg.eos_estimated_gen_length_factor * len(text)))

    @torch.inference_mode()
    def tts_batch(self, texts: List[str], ref_audios, ref_transcripts, cfg: InferenceConfig = InferenceConfig(),
                  seed: int = 0, utt_base: int = 0):
        """Batched tts(): every row equals a single reference call.  Returns a list of (L0 codes, waveform).
        Randomness: Philox streams keyed by (seed, utt_base + row) -- the same (seed, utterance id) reproduces the same audio
        whatever the batch composition or the number of GPUs.  `cfg.use_kv_cache` is accepted for API parity; the engine
        always uses its KV cache (the reference's two paths agree to 1e-6 on the logits)."""
        assert cfg.beam_width == 1, "Only beam size of 1 is currently supported."
        preps = [self._prepare(t, a, r, cfg) for t, a, r in zip(texts, ref_audios, ref_transcripts)]
        eng, tt = self.engine, self.texttok
        max_len = cfg.generate_max_len_override if cfg.generate_max_len_override > 1 else 2000
        eos = len(tt.vocab) + self.speechtok.special_tokens["<|endofspeech|>"]
        results = [None] * len(preps)
        for s in range(0, len(preps), 32):  # the AR kernels keep <= 32 rows in flight
            chunk = preps[s:s + 32]
            acfg = eng.make_ar_cfg(cfg, max_len, eos)
            ids, hit, _ = eng.ar_generate([p["prompt"] for p in chunk], [p["spk_ref"] for p in chunk],
                                          [p["n_phones"] for p in chunk], acfg, seed=seed,
                                          utt_ids=list(range(utt_base + s, utt_base + s + len(chunk))))
            toks = []
            for p, seq, h in zip(chunk, ids, hit):
                if h:
                    logging.warning(f"[autoregressive generation] output length = {len(seq)} -- inference likely failed or input too long!")
                toks.append(np.clip(seq.astype(np.int64) - len(tt.vocab), 0, None)[p["first_codec_idx"]:].tolist())
            # speech BPE -> Encodec L0 codes for the whole chunk at once when the tokeniser offers it (bpe.py)
            dec = (self.speechtok.decode_int_batch(toks) if hasattr(self.speechtok, "decode_int_batch")
                   else [self.speechtok.decode_int(t) for t in toks])
            l0s = [np.asarray([c for c in d if type(c) == int], dtype=np.int32) for d in dec]
            ncfg = eng.make_nar_cfg(cfg, T=self.default_T)
            codes = eng.nar_infer([p["text_tokens"] for p in chunk], [p["spk_ref"] for p in chunk], l0s, ncfg, seed=seed,
                                  utt_ids=list(range(utt_base + s, utt_base + s + len(chunk))))
            outs = []
            for p, c in zip(chunk, codes):
                skip = len(p["spk_ref"]) if cfg.deep_clone else 0  # second crop of inference.py:300-301
                outs.append(c[skip:])
            # vocode final output and trim silences (inference.py:304-305): one call, the frame powers of the trim are
            # computed on the device behind the overlap-add (utterances too short for a 2048-sample frame go to the host path,
            # which raises like the reference does)
            if all(len(o) * eng.dims["voc_hop"] > 1024 for o in outs):
                wavs, bounds = eng.vocode_trim(outs, cfg.trim_db, bandwidth_id=1)
            else:
                wavs = eng.vocode(outs, bandwidth_id=1)
                bounds = trim_bounds_batch(wavs, top_db=cfg.trim_db)
            for i, (l0, w, (a, b)) in enumerate(zip(l0s, wavs, bounds)):
                results[s + i] = (torch.from_numpy(l0.astype(np.int64)).to(self.device), torch.from_numpy(w[a:b].copy()))
        return results

    @torch.inference_mode()
    def tts(self, text: str, ref_audio: torch.Tensor, ref_transcript: Optional[str] = None,
            cfg: Optional[InferenceConfig] = InferenceConfig(), seed: Optional[int] = None):
        """Same contract as the reference's tts() (inference.py:201-307): (AR L0 codes on the device, trimmed 24 kHz
        waveform on the CPU).  Like the reference (which never seeds torch), consecutive calls draw different randomness:
        the Philox seed defaults to torch.initial_seed() and the utterance id to a per-object call counter, so
        torch.manual_seed(s) before the first call makes a run reproducible; pass seed= for an explicit stream."""
        if seed is None:
            seed = torch.initial_seed() & 0x7FFFFFFFFFFFFFFF
        codes, wav = self.tts_batch([text], [ref_audio], [ref_transcript], cfg, seed=seed, utt_base=self._calls)[0]
        self._calls += 1
        return codes, wav
