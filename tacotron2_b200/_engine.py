"""Host-side glue between the nn.Module boundary and libt2b200.so: weight table, handle cache,
workspaces, dropout-mask injection.  All arithmetic happens in the shared library."""
import contextlib
import ctypes as C
import threading

import torch

from . import _capi


def weight_table_spec(hp):
    """(name, shape) of the 84 state_dict entries in the reference's order."""
    E, K = hp.encoder_embedding_dim, hp.encoder_kernel_size
    s = [("embedding.weight", (hp.n_symbols, hp.symbols_embedding_dim))]

    def bn(p, c):
        return [(p + "weight", (c,)), (p + "bias", (c,)), (p + "running_mean", (c,)),
                (p + "running_var", (c,)), (p + "num_batches_tracked", ())]
    for i in range(hp.encoder_n_convolutions):
        p = "encoder.convolutions.%d." % i
        s += [(p + "0.conv.weight", (E, E, K)), (p + "0.conv.bias", (E,))] + bn(p + "1.", E)
    H = E // 2
    for suf in ("", "_reverse"):
        s += [("encoder.lstm.weight_ih_l0" + suf, (4 * H, E)), ("encoder.lstm.weight_hh_l0" + suf, (4 * H, H)),
              ("encoder.lstm.bias_ih_l0" + suf, (4 * H,)), ("encoder.lstm.bias_hh_l0" + suf, (4 * H,))]
    d = "decoder."
    nm = hp.n_mel_channels * hp.n_frames_per_step
    A, D, P = hp.attention_rnn_dim, hp.decoder_rnn_dim, hp.prenet_dim
    s += [(d + "prenet.layers.0.linear_layer.weight", (P, nm)), (d + "prenet.layers.1.linear_layer.weight", (P, P)),
          (d + "attention_rnn.weight_ih", (4 * A, P + E)), (d + "attention_rnn.weight_hh", (4 * A, A)),
          (d + "attention_rnn.bias_ih", (4 * A,)), (d + "attention_rnn.bias_hh", (4 * A,))]
    a = d + "attention_layer."
    s += [(a + "query_layer.linear_layer.weight", (hp.attention_dim, A)),
          (a + "memory_layer.linear_layer.weight", (hp.attention_dim, E)),
          (a + "v.linear_layer.weight", (1, hp.attention_dim)),
          (a + "location_layer.location_conv.conv.weight",
           (hp.attention_location_n_filters, 2, hp.attention_location_kernel_size)),
          (a + "location_layer.location_dense.linear_layer.weight", (hp.attention_dim, hp.attention_location_n_filters))]
    s += [(d + "decoder_rnn.weight_ih", (4 * D, A + E)), (d + "decoder_rnn.weight_hh", (4 * D, D)),
          (d + "decoder_rnn.bias_ih", (4 * D,)), (d + "decoder_rnn.bias_hh", (4 * D,)),
          (d + "linear_projection.linear_layer.weight", (nm, D + E)), (d + "linear_projection.linear_layer.bias", (nm,)),
          (d + "gate_layer.linear_layer.weight", (1, D + E)), (d + "gate_layer.linear_layer.bias", (1,))]
    PD, PK, n = hp.postnet_embedding_dim, hp.postnet_kernel_size, hp.postnet_n_convolutions
    for i in range(n):
        ci = hp.n_mel_channels if i == 0 else PD
        co = hp.n_mel_channels if i == n - 1 else PD
        p = "postnet.convolutions.%d." % i
        s += [(p + "0.conv.weight", (co, ci, PK)), (p + "0.conv.bias", (co,))] + bn(p + "1.", co)
    return s


# ---- dropout mask injection (parity tests feed the SAME Bernoulli masks to oracle and engine) ----
_tls = threading.local()


@contextlib.contextmanager
def dropout_masks(prenet=None, att=None, dec=None, enc=None, post=None):
    """uint8 keep masks (1 = keep).  prenet: (steps, 2, B, 256) [Decoder.forward: steps = T_mel+1];
    att / dec: (T_mel, B, 1024); enc: (3, B, 512, T_text); post: list/tuple of 5 masks in the
    reference layout [(B,512,T)]*4 + [(B,80,T)] (training only).  None => in-kernel Philox."""
    prev = getattr(_tls, "masks", None)
    _tls.masks = dict(prenet=prenet, att=att, dec=dec, enc=enc, post=post)
    try:
        yield
    finally:
        _tls.masks = prev


def current_masks():
    return getattr(_tls, "masks", None) or dict(prenet=None, att=None, dec=None, enc=None, post=None)


# Bumped by code that updates parameters in place without going through torch (the fused optimizer): part of the
# handle-cache key, so the packed device-side copies are rebuilt.
_weights_generation = [0]


def bump_weights_generation():
    _weights_generation[0] += 1


def invalidate_weights():
    """Public: every engine re-packs its device-side weight copies on its next call (use after in-place parameter writes
    that bypass torch's version counter, e.g. ``p.data.copy_()`` for EMA / SWA / weight surgery)."""
    bump_weights_generation()


_seed_counter = [0]


def next_seed():
    """Philox seed for one engine call: torch's global seed + a call counter."""
    _seed_counter[0] += 1
    return (torch.initial_seed() * 1000003 + _seed_counter[0]) & 0xFFFFFFFFFFFFFFFF


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _u8(mask, device):
    if mask is None:
        return None
    return mask.to(device=device, dtype=torch.uint8).contiguous()


class Engine:
    """One T2Model handle (packed weights on one device) + cached workspaces."""

    def __init__(self, hp):
        self.hp = hp
        self.spec = weight_table_spec(hp)
        self.handle = None
        self.key = None
        self.held = None
        self.device = None
        self._ws = {}
        self.impl = _capi.IMPL_AUTO

    # -- weights ---------------------------------------------------------------------------------
    def invalidate(self):
        """Forget the cache key: the next call re-packs every device-side copy from the live parameters."""
        self.key = None

    def ensure(self, named, force=False):
        """named: dict full-name -> tensor (missing entries are replaced by zeros).  The packed copies are rebuilt when the
        key (global generation, per-tensor pointer / torch version counter / dtype) changed or when `force` is set.
        Writes through ``.data`` do not move torch's version counter: callers doing that use
        ``model.invalidate_weights()`` / ``tacotron2_b200.invalidate_weights()``."""
        dev = None
        for t in named.values():
            if t.is_cuda:
                dev = t.device
                break
        if dev is None:
            raise RuntimeError("tacotron2_b200: the model must live on a CUDA device (B200); there is no "
                               "CPU path -- call .cuda() first")
        key = (_weights_generation[0],) + tuple(
            (named[n].data_ptr(), named[n]._version, named[n].dtype) if n in named else None for n, _ in self.spec)
        if self.handle is not None and key == self.key and dev == self.device and not force:
            return
        L = _capi.lib()
        held, ptrs = [], (C.c_void_p * _capi.T2_NUM_WEIGHTS)()
        for i, (n, shape) in enumerate(self.spec):
            t = named.get(n)
            if n.endswith("num_batches_tracked"):
                ptrs[i] = None
                continue
            if t is None:
                t = torch.zeros(shape, device=dev, dtype=torch.float32)
                if n.endswith("running_var"):
                    t.fill_(1.0)
            else:
                if tuple(t.shape) != tuple(shape):
                    raise RuntimeError("tacotron2_b200: %s has shape %s, expected %s" % (n, tuple(t.shape), shape))
                t = t.detach()
                if t.dtype != torch.float32 or not t.is_contiguous():
                    t = t.float().contiguous()
            held.append(t)
            ptrs[i] = t.data_ptr()
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        with torch.cuda.device(dev):
            if self.handle is None or dev != self.device:
                self.close()
                hp = self.hp
                cfg = _capi.T2Config(
                    hp.n_mel_channels, hp.n_symbols, hp.symbols_embedding_dim, hp.encoder_kernel_size,
                    hp.encoder_n_convolutions, hp.encoder_embedding_dim, hp.attention_rnn_dim,
                    hp.decoder_rnn_dim, hp.prenet_dim, hp.attention_dim, hp.attention_location_n_filters,
                    hp.attention_location_kernel_size, hp.postnet_embedding_dim, hp.postnet_kernel_size,
                    hp.postnet_n_convolutions, hp.p_attention_dropout, hp.p_decoder_dropout, 1e-5)
                if hp.n_frames_per_step != 1:
                    raise RuntimeError("n_frames_per_step != 1 is not supported (hparams.py:56)")
                h = C.c_void_p()
                _capi.check(L.t2_model_create(C.byref(h), C.byref(cfg), ptrs, _capi.T2_NUM_WEIGHTS, stream))
                self.handle = h
            else:
                _capi.check(L.t2_model_refresh(self.handle, ptrs, _capi.T2_NUM_WEIGHTS, stream))
        self.key, self.held, self.device = key, held, dev

    def close(self):
        if self.handle is not None:
            _capi.lib().t2_model_destroy(self.handle)
            self.handle = None
            self.key = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _workspace(self, tag, nbytes):
        t = self._ws.get(tag)
        if t is None or t.numel() < nbytes or t.device != self.device:
            t = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
            self._ws[tag] = t
        return t

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # -- encoder ---------------------------------------------------------------------------------
    def encoder(self, text=None, embedded=None, lengths=None, training=False, keep=None, stash=None, seed=None):
        L = _capi.lib()
        src = text if text is not None else embedded
        B, T = int(src.shape[0]), int(src.shape[1])
        memory = torch.empty(B, T, self.hp.encoder_embedding_dim, device=self.device, dtype=torch.float32)
        ws = self._workspace("enc", L.t2_encoder_workspace_bytes(self.handle, B, T))
        a = _capi.T2EncoderArgs()
        if text is not None:
            text = text.to(device=self.device, dtype=torch.int64).contiguous()
            a.text = text.data_ptr()
        else:
            embedded = embedded.to(device=self.device, dtype=torch.float32).contiguous()
            a.embedded = embedded.data_ptr()
        len32 = None
        if lengths is not None:
            len32 = lengths.to(device=self.device, dtype=torch.int32).contiguous()
            a.lengths = len32.data_ptr()
        keep = _u8(keep, self.device)
        a.B, a.T, a.training = B, T, int(bool(training))
        a.keep = keep.data_ptr() if keep is not None else None
        a.seed = next_seed() if seed is None else seed
        a.memory, a.ws, a.ws_bytes = memory.data_ptr(), ws.data_ptr(), ws.numel()
        if stash is not None:
            a.stash, a.stash_bytes = stash.data_ptr(), stash.numel()
        with torch.cuda.device(self.device):
            _capi.check(L.t2_encoder_forward(self.handle, C.byref(a), self._stream()))
        return memory

    def stash_buffer(self, kind, *dims):
        n = getattr(_capi.lib(), "t2_%s_stash_bytes" % kind)(self.handle, *dims)
        return torch.empty(int(n), dtype=torch.uint8, device=self.device)

    def encoder_backward(self, text, embedded, lengths, training, keep, seed, stash, d_memory, want_d_embedded, named_grads):
        L = _capi.lib()
        B, T = int(d_memory.shape[0]), int(d_memory.shape[1])
        f32 = dict(device=self.device, dtype=torch.float32)
        a = _capi.T2EncoderBwdArgs()
        if text is not None:
            text = text.to(device=self.device, dtype=torch.int64).contiguous()
            a.text = text.data_ptr()
        else:
            embedded = embedded.to(**f32).contiguous()
            a.embedded = embedded.data_ptr()
        len32 = None
        if lengths is not None:
            len32 = lengths.to(device=self.device, dtype=torch.int32).contiguous()
            a.lengths = len32.data_ptr()
        keep = _u8(keep, self.device)
        a.B, a.T, a.training = B, T, int(bool(training))
        a.keep = keep.data_ptr() if keep is not None else None
        a.seed = seed
        a.stash, a.stash_bytes = stash.data_ptr(), stash.numel()
        d_memory = d_memory.to(**f32).contiguous()
        a.d_memory = d_memory.data_ptr()
        d_emb = torch.empty(B, T, self.hp.encoder_embedding_dim, **f32) if want_d_embedded else None
        a.d_embedded = d_emb.data_ptr() if d_emb is not None else None
        ptrs = self.grad_table(named_grads)
        a.grads, a.n_grads = ptrs, _capi.T2_NUM_WEIGHTS
        ws = self._workspace("enc_bwd", L.t2_encoder_backward_workspace_bytes(self.handle, B, T))
        a.ws, a.ws_bytes = ws.data_ptr(), ws.numel()
        with torch.cuda.device(self.device):
            _capi.check(L.t2_encoder_backward(self.handle, C.byref(a), self._stream()))
        return d_emb

    def postnet_backward(self, B, T, training, add_residual, keep, seed, stash, d_mel_post, named_grads, wgrad_lengths=None):
        """d_mel_post (B, 80, T) -> d_mel (B, T, 80)."""
        L = _capi.lib()
        f32 = dict(device=self.device, dtype=torch.float32)
        a = _capi.T2PostnetBwdArgs()
        a.B, a.T, a.training, a.add_residual = B, T, int(bool(training)), int(bool(add_residual))
        if keep is not None and isinstance(keep, (list, tuple)):
            keep = torch.cat([k.to(torch.uint8).reshape(-1) for k in keep])
        keep = _u8(keep, self.device)
        a.keep = keep.data_ptr() if keep is not None else None
        a.seed = seed
        wl32 = None
        if wgrad_lengths is not None:
            wl32 = wgrad_lengths.to(device=self.device, dtype=torch.int32).contiguous()
            a.wgrad_lengths = wl32.data_ptr()
        a.stash, a.stash_bytes = stash.data_ptr(), stash.numel()
        d_mel_post = d_mel_post.to(**f32).contiguous()
        a.d_mel_post = d_mel_post.data_ptr()
        d_mel = torch.empty(B, T, self.hp.n_mel_channels, **f32)
        a.d_mel = d_mel.data_ptr()
        ptrs = self.grad_table(named_grads)
        a.grads, a.n_grads = ptrs, _capi.T2_NUM_WEIGHTS
        ws = self._workspace("post_bwd", L.t2_postnet_backward_workspace_bytes(self.handle, B, T))
        a.ws, a.ws_bytes = ws.data_ptr(), ws.numel()
        with torch.cuda.device(self.device):
            _capi.check(L.t2_postnet_backward(self.handle, C.byref(a), self._stream()))
        return d_mel

    # -- decoder ---------------------------------------------------------------------------------
    def decoder(self, memory, mode, n_steps_cap, memory_lengths=None, teacher_prenet=None, training=False,
                prenet_keep=None, att_keep=None, dec_keep=None, gate_threshold=0.5,
                score_mask_value=-float("inf"), impl=None, stash=None, seed=None):
        L = _capi.lib()
        memory = memory.to(device=self.device, dtype=torch.float32).contiguous()
        B, T = int(memory.shape[0]), int(memory.shape[1])
        cap = int(n_steps_cap)
        # zero-initialised: batches of more than 64 rows run as independent 64-row launches that may stop at
        # different steps; frames past a launch's last step stay zero
        mel = torch.zeros(B, cap, self.hp.n_mel_channels, device=self.device, dtype=torch.float32)
        gate = torch.zeros(B, cap, device=self.device, dtype=torch.float32)
        align = torch.zeros(B, cap, T, device=self.device, dtype=torch.float32)
        mel_lengths = torch.zeros(B, device=self.device, dtype=torch.int32)
        n_steps = torch.zeros(1, device=self.device, dtype=torch.int32)
        ws = self._workspace("dec", L.t2_decoder_workspace_bytes(self.handle, B, T, cap))
        a = _capi.T2DecoderArgs()
        a.mode, a.impl, a.training = mode, self.impl if impl is None else impl, int(bool(training))
        a.memory = memory.data_ptr()
        len32 = None
        if memory_lengths is not None:
            len32 = memory_lengths.to(device=self.device, dtype=torch.int32).contiguous()
            a.memory_lengths = len32.data_ptr()
        a.B, a.T_enc, a.n_steps_cap = B, T, cap
        if teacher_prenet is not None:
            teacher_prenet = teacher_prenet.contiguous()
            a.teacher_prenet = teacher_prenet.data_ptr()
        pk, ak, dk = _u8(prenet_keep, self.device), _u8(att_keep, self.device), _u8(dec_keep, self.device)
        a.prenet_keep = pk.data_ptr() if pk is not None else None
        a.att_keep = ak.data_ptr() if ak is not None else None
        a.dec_keep = dk.data_ptr() if dk is not None else None
        a.seed = next_seed() if seed is None else seed
        if stash is not None:
            a.stash, a.stash_bytes = stash.data_ptr(), stash.numel()
        a.gate_threshold = float(gate_threshold)
        a.score_mask_value = float(score_mask_value)
        a.mel, a.gate, a.align = mel.data_ptr(), gate.data_ptr(), align.data_ptr()
        a.mel_lengths, a.n_steps = mel_lengths.data_ptr(), n_steps.data_ptr()
        a.ws, a.ws_bytes = ws.data_ptr(), ws.numel()
        with torch.cuda.device(self.device):
            _capi.check(L.t2_decoder_run(self.handle, C.byref(a), self._stream()))
        self._last_decoder_args = (a, memory, len32, teacher_prenet, pk, ak, dk, ws, mel, gate, align, mel_lengths, n_steps)
        return mel, gate, align, mel_lengths, n_steps

    def decoder_profile(self):
        """Per-phase SM cycles of the last persistent decoder run: dict phase -> [cta0, cta60, cta100]."""
        a = self._last_decoder_args[0]
        out = (C.c_int64 * 72)()
        _capi.check(_capi.lib().t2_decoder_profile(C.byref(a), out))
        names = ["E0 x2->att gemm", "E0 epilogue(ah)", "B1", "E1 ah->dec/att/q gemm", "B2", "attention", "B3",
                 "E2 ctx gemm", "E2 epilogue(dh)", "B4", "E3 dh gemm", "E3 epilogue(mel/x1)", "B5", "E4 x1 gemm+epi", "B6",
                 "att:im2col", "att:mma", "att:energies", "att:softmax", "att:context"]
        return {names[i]: [int(out[s * 24 + i]) for s in range(3)] for i in range(len(names))}

    def grad_table(self, named_grads):
        """(ctypes pointer array of T2_NUM_WEIGHTS entries, held tensors): named_grads maps full state_dict names to
        the contiguous fp32 tensors the library overwrites with that parameter's gradient."""
        ptrs = (C.c_void_p * _capi.T2_NUM_WEIGHTS)()
        for i, (n, _) in enumerate(self.spec):
            t = named_grads.get(n)
            ptrs[i] = t.data_ptr() if t is not None else None
        return ptrs

    def decoder_stash(self, B, T_enc, T_mel):
        return self.stash_buffer("decoder", B, T_enc, T_mel)

    def decoder_backward(self, memory, memory_lengths, teacher_prenet, align, stash, seed, training, att_keep, dec_keep,
                         score_mask_value, d_mel, d_gate, d_align, named_grads):
        """Backward of the teacher-forced decoder run that filled `stash`.  d_mel (B, T, 80), d_gate (B, T), d_align
        (B, T, T_enc) or None.  Returns (d_memory (B, T_enc, 512), d_prenet (T, B, 256))."""
        L = _capi.lib()
        B, Te = int(memory.shape[0]), int(memory.shape[1])
        T = int(align.shape[1])
        f32 = dict(device=self.device, dtype=torch.float32)
        d_memory = torch.empty(B, Te, self.hp.encoder_embedding_dim, **f32)
        d_prenet = torch.empty(T, B, self.hp.prenet_dim, **f32)
        ws = self._workspace("dec_bwd", L.t2_decoder_backward_workspace_bytes(self.handle, B, Te, T))
        a = _capi.T2DecoderBwdArgs()
        memory = memory.contiguous()
        a.memory = memory.data_ptr()
        len32 = None
        if memory_lengths is not None:
            len32 = memory_lengths.to(device=self.device, dtype=torch.int32).contiguous()
            a.memory_lengths = len32.data_ptr()
        a.B, a.T_enc, a.T_mel, a.training = B, Te, T, int(bool(training))
        a.teacher_prenet = teacher_prenet.data_ptr()
        ak, dk = _u8(att_keep, self.device), _u8(dec_keep, self.device)
        a.att_keep = ak.data_ptr() if ak is not None else None
        a.dec_keep = dk.data_ptr() if dk is not None else None
        a.seed, a.score_mask_value = seed, float(score_mask_value)
        a.align, a.stash, a.stash_bytes = align.data_ptr(), stash.data_ptr(), stash.numel()
        d_mel, d_gate = d_mel.to(**f32).contiguous(), d_gate.to(**f32).contiguous()
        a.d_mel, a.d_gate = d_mel.data_ptr(), d_gate.data_ptr()
        if d_align is not None:
            d_align = d_align.to(**f32).contiguous()
            a.d_align = d_align.data_ptr()
        a.d_memory, a.d_prenet = d_memory.data_ptr(), d_prenet.data_ptr()
        ptrs = self.grad_table(named_grads)
        a.grads, a.n_grads = ptrs, _capi.T2_NUM_WEIGHTS
        a.ws, a.ws_bytes = ws.data_ptr(), ws.numel()
        with torch.cuda.device(self.device):
            _capi.check(L.t2_decoder_backward(self.handle, C.byref(a), self._stream()))
        return d_memory, d_prenet

    def prenet_backward(self, frames, keep, seed, d_out, named_grads):
        L = _capi.lib()
        frames = frames.to(device=self.device, dtype=torch.float32).contiguous()
        M = int(frames.shape[0])
        d_out = d_out.contiguous()
        ws = self._workspace("pre_bwd", L.t2_prenet_backward_workspace_bytes(self.handle, M))
        keep = _u8(keep, self.device)
        a = _capi.T2PrenetBwdArgs()
        a.frames, a.M = frames.data_ptr(), M
        a.keep = keep.data_ptr() if keep is not None else None
        a.seed, a.d_out = seed, d_out.data_ptr()
        ptrs = self.grad_table(named_grads)
        a.grads, a.n_grads = ptrs, _capi.T2_NUM_WEIGHTS
        a.ws, a.ws_bytes = ws.data_ptr(), ws.numel()
        with torch.cuda.device(self.device):
            _capi.check(L.t2_prenet_backward(self.handle, C.byref(a), self._stream()))

    def prenet(self, frames, keep=None, seed=None):
        """frames (M, 80) -> (M, 256); keep (2, M, 256) uint8 or None."""
        L = _capi.lib()
        frames = frames.to(device=self.device, dtype=torch.float32).contiguous()
        M = int(frames.shape[0])
        out = torch.empty(M, self.hp.prenet_dim, device=self.device, dtype=torch.float32)
        ws = self._workspace("pre", M * self.hp.prenet_dim * 4)
        keep = _u8(keep, self.device)
        with torch.cuda.device(self.device):
            _capi.check(L.t2_prenet_forward(self.handle, frames.data_ptr(), M,
                                            keep.data_ptr() if keep is not None else None,
                                            next_seed() if seed is None else seed,
                                            out.data_ptr(), ws.data_ptr(), ws.numel(), self._stream()))
        return out

    # -- postnet ---------------------------------------------------------------------------------
    def postnet(self, mel_btc, lengths=None, add_residual=True, training=False, keep=None, stash=None, seed=None):
        """mel_btc: (B, T, 80) time-major rows (batch stride may exceed T*80).  Returns (B, 80, T)."""
        L = _capi.lib()
        assert mel_btc.dtype == torch.float32 and mel_btc.stride(2) == 1 and mel_btc.stride(1) == mel_btc.shape[2]
        B, T = int(mel_btc.shape[0]), int(mel_btc.shape[1])
        out = torch.empty(B, self.hp.n_mel_channels, T, device=self.device, dtype=torch.float32)
        ws = self._workspace("post", L.t2_postnet_workspace_bytes(self.handle, B, T))
        a = _capi.T2PostnetArgs()
        a.mel, a.mel_batch_stride = mel_btc.data_ptr(), int(mel_btc.stride(0))
        len32 = None
        if lengths is not None:
            len32 = lengths.to(device=self.device, dtype=torch.int32).contiguous()
            a.lengths = len32.data_ptr()
        a.B, a.T, a.training = B, T, int(bool(training))
        if keep is not None and isinstance(keep, (list, tuple)):
            keep = torch.cat([k.to(torch.uint8).reshape(-1) for k in keep])
        keep = _u8(keep, self.device)
        a.keep = keep.data_ptr() if keep is not None else None
        a.seed = next_seed() if seed is None else seed
        a.add_residual = int(bool(add_residual))
        a.mel_post, a.ws, a.ws_bytes = out.data_ptr(), ws.data_ptr(), ws.numel()
        if stash is not None:
            a.stash, a.stash_bytes = stash.data_ptr(), stash.numel()
        with torch.cuda.device(self.device):
            _capi.check(L.t2_postnet_forward(self.handle, C.byref(a), self._stream()))
        return out

    # -- end to end with host buffers (bench e2e leg) ------------------------------------------------
    def infer_host(self, text_host, max_steps, gate_threshold=0.5, impl=None, out_host=None):
        """text_host: pinned int64 (B, T).  Returns (mel_post_host (B,80,max_steps), lengths, n_steps)."""
        L = _capi.lib()
        B, T = int(text_host.shape[0]), int(text_host.shape[1])
        ws = self._workspace("e2e", L.t2_infer_workspace_bytes(self.handle, B, T, max_steps))
        if out_host is None:
            out_host = (torch.empty(B, self.hp.n_mel_channels, max_steps, dtype=torch.float32).pin_memory(),
                        torch.empty(B, dtype=torch.int32).pin_memory(), torch.empty(1, dtype=torch.int32).pin_memory())
        mel, lens, ns = out_host
        with torch.cuda.device(self.device):
            _capi.check(L.t2_infer_host(self.handle, text_host.data_ptr(), B, T, int(max_steps), float(gate_threshold),
                                        next_seed(), self.impl if impl is None else impl, mel.data_ptr(),
                                        lens.data_ptr(), ns.data_ptr(), ws.data_ptr(), ws.numel(), self._stream()))
        return mel, lens, ns
