"""CPU oracle: a numpy restatement of the rwkv.cpp evaluation path (RWKV v4 / v5.1 / v5.2 / v6 / v7).

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE. Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import it; the product (rwkv.cpp_b200/) never does and has no CPU fallback.

Parity status: PINNED. tests/test_oracle.py checks this file against
  * the reference's golden vectors tests/expected-logits-*.bin (copied to tests/golden/logits/),
  * the known-answer tables of tests/test_tiny_rwkv.c:38-54,70-134, and
  * outputs of the unmodified reference compiled by oracle/Makefile (oracle/_ref/), committed as
    tests/golden/ref_outputs.npz by tests/golden/make_ref_outputs.py.

Every function cites the reference file:line it restates (paths relative to /root/reference).
Arithmetic follows the reference's x86 CPU path: quantised weights are multiplied with activations
that were first quantised to Q8_0 / Q8_1 (ggml-cpu.c:265-306 `vec_dot_type`), F16 weights with
activations rounded to fp16 (ggml-cpu.c:259-264), all accumulation in fp32.
"""
import numpy as np

from ggml_file import (TYPE_FP16, TYPE_FP32, TYPE_Q4_0, TYPE_Q4_1, TYPE_Q5_0, TYPE_Q5_1, TYPE_Q8_0,
                       read_model_file)

F32 = np.float32


# --------------------------------------------------------------------------------------------
# Quant block (de)coding -- ggml-common.h:161-221, ggml-quants.c:31-363
# --------------------------------------------------------------------------------------------

def _f16_field(raw, nblocks, bsize, off):
    """fp16 field at byte offset `off` of each block, as float32."""
    b = raw.reshape(nblocks, bsize)
    return np.ascontiguousarray(b[:, off:off + 2]).view(np.float16).reshape(nblocks).astype(F32)


def _nibbles(qs):
    """ggml-quants.c:262-275: element j = low nibble of qs[j], element j+16 = high nibble."""
    return np.concatenate([qs & 0x0F, qs >> 4], axis=1).astype(np.int32)


def _fifth_bits(qh_bytes):
    """ggml-quants.c:312-320: bit j of the little-endian u32 qh is the 5th bit of element j."""
    qh = np.ascontiguousarray(qh_bytes).view("<u4").reshape(-1, 1)
    return ((qh >> np.arange(32, dtype=np.uint32)) & 1).astype(np.int32)


def unpack_blocks(dtype, raw, n_elements):
    """Return (q:int32[nb,32], d:f32[nb], m:f32[nb] or None, offset:int) with w = (q - offset)*d (+ m)."""
    nb = n_elements // 32
    raw = np.ascontiguousarray(raw).reshape(-1)
    if dtype == TYPE_Q4_0:
        b = raw.reshape(nb, 18)
        return _nibbles(b[:, 2:18]), _f16_field(raw, nb, 18, 0), None, 8
    if dtype == TYPE_Q4_1:
        b = raw.reshape(nb, 20)
        return _nibbles(b[:, 4:20]), _f16_field(raw, nb, 20, 0), _f16_field(raw, nb, 20, 2), 0
    if dtype == TYPE_Q5_0:
        b = raw.reshape(nb, 22)
        return _nibbles(b[:, 6:22]) | (_fifth_bits(b[:, 2:6]) << 4), _f16_field(raw, nb, 22, 0), None, 16
    if dtype == TYPE_Q5_1:
        b = raw.reshape(nb, 24)
        return (_nibbles(b[:, 8:24]) | (_fifth_bits(b[:, 4:8]) << 4), _f16_field(raw, nb, 24, 0),
                _f16_field(raw, nb, 24, 2), 0)
    if dtype == TYPE_Q8_0:
        b = raw.reshape(nb, 34)
        return b[:, 2:34].view(np.int8).astype(np.int32), _f16_field(raw, nb, 34, 0), None, 0
    raise ValueError("not a quantized type: %r" % dtype)


def dequantize(dtype, raw, n_elements):
    """dequantize_row_q* (ggml-quants.c:255-363) / fp16 / fp32 -> float32[n_elements]."""
    if dtype == TYPE_FP32:
        return np.ascontiguousarray(raw).view(F32).reshape(n_elements).copy()
    if dtype == TYPE_FP16:
        return np.ascontiguousarray(raw).view(np.float16).reshape(n_elements).astype(F32)
    q, d, m, off = unpack_blocks(dtype, raw, n_elements)
    w = (q - off).astype(F32) * d[:, None]
    if m is not None:
        w = w + m[:, None]
    return w.reshape(n_elements)


def _trunc_i8(x):
    # C `(int8_t)(float)` for in-range values: truncation toward zero
    return np.trunc(x).astype(np.int32)


def quantize_row_ref(dtype, x):
    """quantize_row_q{4_0,4_1,5_0,5_1,8_0}_ref (ggml-quants.c:31-217): float32[k] -> raw block bytes."""
    x = np.asarray(x, dtype=F32).reshape(-1, 32)
    nb = x.shape[0]
    if dtype in (TYPE_Q4_0, TYPE_Q5_0):
        levels = 8 if dtype == TYPE_Q4_0 else 16
        idx = np.argmax(np.abs(x), axis=1)  # first maximum, as `if (amax < fabsf(v))` (ggml-quants.c:45-48)
        mx = x[np.arange(nb), idx]
        d = (mx / F32(-levels)).astype(F32)
        with np.errstate(divide="ignore"):
            idv = np.where(d != 0, F32(1.0) / d, F32(0)).astype(F32)
        q = np.minimum(2 * levels - 1, _trunc_i8(x * idv[:, None] + F32(levels + 0.5))).astype(np.uint8)
        dm = d.astype(np.float16).view(np.uint8).reshape(nb, 2)
    elif dtype in (TYPE_Q4_1, TYPE_Q5_1):
        nlev = 15 if dtype == TYPE_Q4_1 else 31
        mn, mx = x.min(axis=1), x.max(axis=1)
        d = ((mx - mn) / F32(nlev)).astype(F32)
        with np.errstate(divide="ignore"):
            idv = np.where(d != 0, F32(1.0) / d, F32(0)).astype(F32)
        v = (x - mn[:, None]) * idv[:, None] + F32(0.5)
        if dtype == TYPE_Q4_1:
            q = np.minimum(15, _trunc_i8(v)).astype(np.uint8)
        else:
            q = np.trunc(v).astype(np.uint8)  # (uint8_t)(x0 + 0.5f), ggml-quants.c:176
        dm = np.stack([d.astype(np.float16), mn.astype(np.float16)], axis=1).view(np.uint8).reshape(nb, 4)
    elif dtype == TYPE_Q8_0:
        amax = np.abs(x).max(axis=1)
        d = (amax / F32(127)).astype(F32)
        with np.errstate(divide="ignore"):
            idv = np.where(d != 0, F32(1.0) / d, F32(0)).astype(F32)
        v = x * idv[:, None]
        q = (np.sign(v) * np.floor(np.abs(v) + F32(0.5))).astype(np.int8)  # roundf: half away from zero
        return np.concatenate([d.astype(np.float16).view(np.uint8).reshape(nb, 2), q.view(np.uint8)], axis=1).reshape(-1)
    else:
        raise ValueError(dtype)
    qs = (q[:, :16] & 0x0F) | ((q[:, 16:] & 0x0F) << 4)
    if dtype in (TYPE_Q5_0, TYPE_Q5_1):
        bits = ((q >> 4) & 1).astype(np.uint32)
        qh = (bits << np.arange(32, dtype=np.uint32)).sum(axis=1, dtype=np.uint32)
        qh = qh.astype("<u4").view(np.uint8).reshape(nb, 4)
        return np.concatenate([dm, qh, qs], axis=1).reshape(-1)
    return np.concatenate([dm, qs], axis=1).reshape(-1)


def quantize_activations(x, with_sum):
    """quantize_row_q8_0 / quantize_row_q8_1, x86 path (ggml-cpu-quants.c:781-846, 1085-1160).

    x: float32[K, T]. Returns q int32[K/32, 32, T], d f32[K/32, T] (fp16-rounded), s f32[K/32, T] or None.
    d = fp16(amax/127); q = rint(x * (127/amax)) (round-half-even, _mm256_round_ps NEAREST);
    s = fp16(d_f32 * sum(q)).
    """
    K, T = x.shape
    xb = x.reshape(K // 32, 32, T)
    amax = np.abs(xb).max(axis=1)
    d32 = (amax / F32(127)).astype(F32)
    with np.errstate(divide="ignore", invalid="ignore"):
        idv = np.where(amax != 0, F32(127) / amax, F32(0)).astype(F32)
    q = np.rint(xb * idv[:, None, :]).astype(np.int32)
    d = d32.astype(np.float16).astype(F32)
    s = None
    if with_sum:
        s = (d32 * q.sum(axis=1).astype(F32)).astype(np.float16).astype(F32)
    return q, d, s


class Mat:
    """A 2-D weight consumed through ggml_mul_mat: y[M,T] = W[M,K] . x[K,T] (ggml ne = [K, M])."""

    def __init__(self, tensor):
        self.dtype = tensor.dtype
        ne = tensor.ne3
        self.K, self.M = ne[0], ne[1] * ne[2]
        n = self.K * self.M
        if self.dtype in (TYPE_FP32, TYPE_FP16):
            self.w = dequantize(self.dtype, tensor.raw, n).reshape(self.M, self.K)
        else:
            q, d, m, off = unpack_blocks(self.dtype, tensor.raw, n)
            nbk = self.K // 32
            self.q = (q - off).reshape(self.M, nbk, 32)
            self.d = d.reshape(self.M, nbk)
            self.m = None if m is None else m.reshape(self.M, nbk)

    def dense(self):
        if self.dtype in (TYPE_FP32, TYPE_FP16):
            return self.w
        w = self.q.astype(F32) * self.d[:, :, None]
        if self.m is not None:
            w = w + self.m[:, :, None]
        return w.reshape(self.M, self.K)

    def mul(self, x):
        """ggml_compute_forward_mul_mat (ggml-cpu.c:7377) with the per-type vec_dot:
        q4_0/q5_0/q8_0 x q8_0 (ggml-cpu-quants.c:2302-2316, 2944-2963, 3690-3698): sum_b (dW*dA) * isum
        q4_1/q5_1 x q8_1 (ggml-cpu-quants.c:2594-2609, 3318-3338): sum_b (dW*dA) * isum + mW * sA
        f16 (ggml-cpu.c:1463): activations rounded to fp16, fp32 accumulate.  f32: plain."""
        x = np.asarray(x, dtype=F32)
        squeeze = x.ndim == 1
        if squeeze:
            x = x[:, None]
        assert x.shape[0] == self.K, (x.shape, self.K)
        if self.dtype == TYPE_FP32:
            y = self.w @ x
        elif self.dtype == TYPE_FP16:
            y = self.w @ x.astype(np.float16).astype(F32)
        else:
            qa, da, sa = quantize_activations(x, with_sum=self.m is not None)
            isum = np.einsum("mbj,bjt->mbt", self.q, qa, optimize=True).astype(F32)
            y = (isum * self.d[:, :, None] * da[None, :, :]).sum(axis=1, dtype=F32)
            if self.m is not None:
                y = y + np.einsum("mb,bt->mt", self.m, sa).astype(F32)
        y = y.astype(F32)
        return y[:, 0] if squeeze else y


# --------------------------------------------------------------------------------------------
# Elementwise ops
# --------------------------------------------------------------------------------------------

def norm(x, eps):
    """ggml_compute_forward_norm_f32 (ggml-cpu.c:6880-6929) over axis 0; sums in double."""
    x = x.astype(F32)
    n = x.shape[0]
    mean = (x.astype(np.float64).sum(axis=0) / n).astype(F32)
    v = x - mean
    var = ((v * v).astype(np.float64).sum(axis=0) / n).astype(F32)
    scale = F32(1.0) / np.sqrt(var + F32(eps), dtype=F32)
    return (v * scale).astype(F32)


def layer_norm(x, w, b):
    """rwkv_layer_norm (rwkv_operators.inc:93-97): ggml_norm(eps 1e-5) * w + b."""
    if x.ndim == 2:
        return norm(x, 1e-5) * w[:, None] + b[:, None]
    return norm(x, 1e-5) * w + b


def sigmoid(x):
    return (F32(1.0) / (F32(1.0) + np.exp(-x, dtype=F32))).astype(F32)


def silu(x):
    return (x / (F32(1.0) + np.exp(-x, dtype=F32))).astype(F32)


def group_norm(x, H, eps):
    """per-head ggml_norm after reshape to [S, H, T] (rwkv_graph.inc:281-284, 376-379)."""
    C, T = x.shape
    return norm(x.reshape(H, C // H, T).transpose(1, 0, 2), eps).transpose(1, 0, 2).reshape(C, T)


# --------------------------------------------------------------------------------------------
# Model
# --------------------------------------------------------------------------------------------

class OracleModel:
    """Restates rwkv_load_model_from_file (rwkv_model_loading.inc:288-419) + graph builders (rwkv_graph.inc)."""

    def __init__(self, path):
        f = read_model_file(path)
        self.file = f
        self.n_vocab, self.n_embed, self.n_layer = f.n_vocab, f.n_embed, f.n_layer
        t = f.tensors
        # arch detection by key presence: rwkv_model_loading.inc:319-340
        self.major, self.minor = 4, 0
        if "blocks.0.att.ln_x.weight" in t:
            self.major, self.minor = 5, (2 if "blocks.0.att.gate.weight" in t else 1)
        if "blocks.0.att.time_maa_x" in t:
            self.major, self.minor = 6, 0
        if "blocks.0.att.r_k" in t:
            self.major, self.minor = 7, 0
        C = self.n_embed
        # rwkv_model_loading.inc:403-409
        if self.major == 7:
            self.head_count = t["blocks.0.att.r_k"].ne3[1]
            self.head_size = C // self.head_count
        elif self.major >= 5:
            self.head_count = t["blocks.0.att.time_decay"].ne3[2]
            self.head_size = C // self.head_count
        else:
            self.head_count, self.head_size = 0, 0
        self._cache = {}

    # parameter access ---------------------------------------------------------------------
    def vec(self, name):
        key = ("v", name)
        if key not in self._cache:
            tt = self.file.tensors[name]
            n = int(np.prod(tt.ne))
            self._cache[key] = dequantize(tt.dtype, tt.raw, n)
        return self._cache[key]

    def mat(self, name):
        key = ("m", name)
        if key not in self._cache:
            self._cache[key] = Mat(self.file.tensors[name])
        return self._cache[key]

    # state ----------------------------------------------------------------------------------
    @property
    def state_len(self):
        """rwkv_get_state_len (rwkv.cpp:171-179)."""
        per = (2 + self.head_size) if self.major >= 5 else 5
        return self.n_embed * per * self.n_layer

    def init_state(self):
        """rwkv_init_state (rwkv_eval.inc:224-241): zeros; v4 pp slot = -1e30."""
        s = np.zeros(self.state_len, dtype=F32)
        if self.major < 5:
            C = self.n_embed
            s.reshape(self.n_layer, 5, C)[:, 4, :] = F32(-1e30)
        return s

    # forward --------------------------------------------------------------------------------
    def eval_sequence(self, tokens, state_in=None, want_logits=True):
        """rwkv_eval / rwkv_eval_sequence (rwkv_eval.inc:38,79) through rwkv_build_{serial,sequential}_graph
        (rwkv_graph.inc:611,744). Returns (logits or None, state_out)."""
        tokens = np.asarray(tokens, dtype=np.int64).reshape(-1)
        T, C, L = len(tokens), self.n_embed, self.n_layer
        state = self.init_state() if state_in is None else np.array(state_in, dtype=F32, copy=True)
        per = (2 + self.head_size) if self.major >= 5 else 5
        st = state.reshape(L, per * C)

        emb = self.file.tensors["emb.weight"]
        emb_w = self.vec("emb.weight").reshape(self.n_vocab, C)
        x = emb_w[tokens].T.astype(F32)  # [C, T]   ggml_get_rows (rwkv_graph.inc:655)
        x = layer_norm(x, self.vec("blocks.0.ln0.weight"), self.vec("blocks.0.ln0.bias"))

        v_first = None
        for i in range(L):
            p = "blocks.%d." % i
            ls = st[i]
            if self.major == 4:
                x = x + self._att_v4(p, x, ls)
                x = x + self._ffn_v4_v5(p, x, ls)
            elif self.major == 5:
                x = x + self._att_v5(p, x, ls)
                x = x + self._ffn_v4_v5(p, x, ls)
            elif self.major == 6:
                x = x + self._att_v6(p, x, ls)
                x = x + self._ffn_v6(p, x, ls)
            else:
                a, v_first = self._att_v7(p, x, ls, v_first, i)
                x = x + a
                x = x + self._ffn_v7(p, x, ls)
            x = x.astype(F32)

        logits = None
        if want_logits:
            xl = layer_norm(x[:, T - 1], self.vec("ln_out.weight"), self.vec("ln_out.bias"))
            logits = self.mat("head.weight").mul(xl)  # rwkv_graph.inc:705-708, 851-854
        return logits, state

    # token shift: rwkv_carry_x (rwkv_graph.inc:56-82)
    def _carry(self, p, ln, x, ls, slot):
        C = self.n_embed
        xx = layer_norm(x, self.vec(p + ln + ".weight"), self.vec(p + ln + ".bias"))
        prev = np.concatenate([ls[slot * C:(slot + 1) * C][:, None], xx[:, :-1]], axis=1)
        ls[slot * C:(slot + 1) * C] = xx[:, -1]
        return xx, prev

    @staticmethod
    def _mix(x, prev, m):
        # x*m + (prev - prev*m): rwkv_graph.inc:94-97
        return x * m[:, None] + (prev - prev * m[:, None])

    def _att_v4(self, p, x, ls):
        """rwkv_att_rkv_v4 / rwkv_att_wkv_v4 / rwkv_att_v4 (rwkv_graph.inc:84-197)."""
        C = self.n_embed
        T = x.shape[1]
        xx, prev = self._carry(p, "ln1", x, ls, 1)
        xk = self._mix(xx, prev, self.vec(p + "att.time_mix_k"))
        xv = self._mix(xx, prev, self.vec(p + "att.time_mix_v"))
        xr = self._mix(xx, prev, self.vec(p + "att.time_mix_r"))
        r = sigmoid(self.mat(p + "att.receptance.weight").mul(xr))
        k = self.mat(p + "att.key.weight").mul(xk)
        v = self.mat(p + "att.value.weight").mul(xv)
        tf, td = self.vec(p + "att.time_first"), self.vec(p + "att.time_decay")
        aa, bb, pp = ls[2 * C:3 * C].copy(), ls[3 * C:4 * C].copy(), ls[4 * C:5 * C].copy()
        wkv = np.empty((C, T), dtype=F32)
        for t in range(T):
            kt, vt = k[:, t], v[:, t]
            ww = tf + kt
            qq = np.maximum(pp, ww)
            e1, e2 = np.exp(pp - qq, dtype=F32), np.exp(ww - qq, dtype=F32)
            wkv[:, t] = (e1 * aa + e2 * vt) / (e1 * bb + e2)
            ww = pp + td
            qq = np.maximum(ww, kt)
            e1, e2 = np.exp(ww - qq, dtype=F32), np.exp(kt - qq, dtype=F32)
            aa, bb, pp = e1 * aa + e2 * vt, e1 * bb + e2, qq
        ls[2 * C:3 * C], ls[3 * C:4 * C], ls[4 * C:5 * C] = aa, bb, pp
        return self.mat(p + "att.output.weight").mul(r * wkv)

    def _ffn_v4_v5(self, p, x, ls):
        """rwkv_ffn_v4_v5 (rwkv_graph.inc:484-511)."""
        xx, prev = self._carry(p, "ln2", x, ls, 0)
        xk = self._mix(xx, prev, self.vec(p + "ffn.time_mix_k"))
        xr = self._mix(xx, prev, self.vec(p + "ffn.time_mix_r"))
        r = sigmoid(self.mat(p + "ffn.receptance.weight").mul(xr))
        k = np.square(np.maximum(self.mat(p + "ffn.key.weight").mul(xk), 0))
        return r * self.mat(p + "ffn.value.weight").mul(k)

    def _wkv6(self, ls, r, k, v, tf, td):
        """ggml_compute_forward_rwkv_wkv6_f32 (ggml-cpu.c:11803-11935). r,k,v [C,T]; tf [C]; td [C,T] or [C].
        state S[h][i_key][j_val] lives at ls[2C:]."""
        C, T = r.shape
        H, S = self.head_count, self.head_size
        st = ls[2 * C:].reshape(H, S, S).copy()
        out = np.zeros((C, T), dtype=F32)
        tfh = tf.reshape(H, S)
        for t in range(T):
            kh, vh, rh = k[:, t].reshape(H, S), v[:, t].reshape(H, S), r[:, t].reshape(H, S)
            tdh = (td[:, t] if td.ndim == 2 else td).reshape(H, S)
            kv = kh[:, :, None] * vh[:, None, :]  # [H, i, j]
            temp = kv * tfh[:, :, None] + st
            # dst_j += temp_ij * r_i, sequential over i in fp32
            acc = np.zeros((H, S), dtype=F32)
            for i in range(S):
                acc = acc + temp[:, i, :] * rh[:, i:i + 1]
            out[:, t] = acc.reshape(C)
            st = (st * tdh[:, :, None] + kv).astype(F32)
        ls[2 * C:] = st.reshape(-1)
        return out

    def _att_v5(self, p, x, ls):
        """rwkv_att_v5 (rwkv_graph.inc:199-292)."""
        C = self.n_embed
        H, S = self.head_count, self.head_size
        xx, prev = self._carry(p, "ln1", x, ls, 1)
        xk = self._mix(xx, prev, self.vec(p + "att.time_mix_k"))
        xv = self._mix(xx, prev, self.vec(p + "att.time_mix_v"))
        xr = self._mix(xx, prev, self.vec(p + "att.time_mix_r"))
        r = self.mat(p + "att.receptance.weight").mul(xr)
        k = self.mat(p + "att.key.weight").mul(xk)
        v = self.mat(p + "att.value.weight").mul(xv)
        g = None
        if self.minor >= 2:
            xg = self._mix(xx, prev, self.vec(p + "att.time_mix_g"))
            g = silu(self.mat(p + "att.gate.weight").mul(xg))
            tf = self.vec(p + "att.time_faaaa")  # [H*S]
            td = self.vec(p + "att.time_decay")
        else:  # v5.1: one value per head, ggml_repeat to [1,S,H] (rwkv_graph.inc:262-267)
            tf = np.repeat(self.vec(p + "att.time_first"), S)
            td = np.repeat(self.vec(p + "att.time_decay"), S)
        y = self._wkv6(ls, r, k, v, tf, td)
        y = group_norm(y, H, 1e-5)
        y = y * self.vec(p + "att.ln_x.weight")[:, None] + self.vec(p + "att.ln_x.bias")[:, None]
        if g is not None:
            y = y * g
        return self.mat(p + "att.output.weight").mul(y)

    def _att_v6(self, p, x, ls):
        """rwkv_att_v6 (rwkv_graph.inc:294-385)."""
        C = self.n_embed
        T = x.shape[1]
        H, S = self.head_count, self.head_size
        xx, prev = self._carry(p, "ln1", x, ls, 1)
        sx = prev - xx
        xxx = sx * self.vec(p + "att.time_maa_x")[:, None] + xx
        z = np.tanh(self.mat(p + "att.time_maa_w1").mul(xxx)).astype(F32)  # [5*mix, T]
        w2t = self.file.tensors[p + "att.time_maa_w2"]  # ne = [mix, C, 5]
        mix = w2t.ne[0]
        w2 = self.vec(p + "att.time_maa_w2").reshape(5, C, mix)
        z = z.reshape(5, mix, T)
        m = np.einsum("icm,imt->ict", w2, z).astype(F32)  # [5, C, T] order w,k,v,r,g
        names = ["w", "k", "v", "r", "g"]
        xs = {}
        for j, nm in enumerate(names):
            xs[nm] = (m[j] + self.vec(p + "att.time_maa_" + nm)[:, None]) * sx + xx
        r = self.mat(p + "att.receptance.weight").mul(xs["r"])
        k = self.mat(p + "att.key.weight").mul(xs["k"])
        v = self.mat(p + "att.value.weight").mul(xs["v"])
        g = silu(self.mat(p + "att.gate.weight").mul(xs["g"]))
        w = self.mat(p + "att.time_decay_w2").mul(np.tanh(self.mat(p + "att.time_decay_w1").mul(xs["w"])).astype(F32))
        w = w + self.vec(p + "att.time_decay")[:, None]
        w = np.exp(-np.exp(w, dtype=F32), dtype=F32)
        y = self._wkv6(ls, r, k, v, self.vec(p + "att.time_faaaa"), w)
        y = group_norm(y, H, 64e-5)
        y = y * self.vec(p + "att.ln_x.weight")[:, None] + self.vec(p + "att.ln_x.bias")[:, None]
        y = y * g
        return self.mat(p + "att.output.weight").mul(y)

    def _ffn_v6(self, p, x, ls):
        """rwkv_ffn_v6 (rwkv_graph.inc:513-531)."""
        xx, prev = self._carry(p, "ln2", x, ls, 0)
        sx = prev - xx
        xk = sx * self.vec(p + "ffn.time_maa_k")[:, None] + xx
        xr = sx * self.vec(p + "ffn.time_maa_r")[:, None] + xx
        r = sigmoid(self.mat(p + "ffn.receptance.weight").mul(xr))
        k = np.square(np.maximum(self.mat(p + "ffn.key.weight").mul(xk), 0))
        return r * self.mat(p + "ffn.value.weight").mul(k)

    def _wkv7(self, ls, r, w, k, v, a, b):
        """rwkv_wkv_v7_impl (rwkv_operators_wkv_v7.inc:37-106). state S[h][i_val][j_key]."""
        C, T = r.shape
        H, S = self.head_count, self.head_size
        st = ls[2 * C:].reshape(H, S, S).copy()
        out = np.zeros((C, T), dtype=F32)
        for t in range(T):
            rh, wh, kh, vh, ah, bh = (z[:, t].reshape(H, S) for z in (r, w, k, v, a, b))
            sa = np.zeros((H, S), dtype=F32)
            for j in range(S):  # sequential fp32 sum over j
                sa = sa + ah[:, j:j + 1] * st[:, :, j]
            st = (st * wh[:, None, :] + vh[:, :, None] * kh[:, None, :] + sa[:, :, None] * bh[:, None, :]).astype(F32)
            y = np.zeros((H, S), dtype=F32)
            for j in range(S):
                y = y + st[:, :, j] * rh[:, j:j + 1]
            out[:, t] = y.reshape(C)
        ls[2 * C:] = st.reshape(-1)
        return out

    def _att_v7(self, p, x, ls, v_first, layer_idx):
        """rwkv_att_v7 (rwkv_graph.inc:387-482)."""
        C = self.n_embed
        T = x.shape[1]
        H, S = self.head_count, self.head_size
        xx, prev = self._carry(p, "ln1", x, ls, 1)
        sx = prev - xx
        mixes = self.vec(p + "att.x_rwkvag").reshape(6, C)  # r, w, k, v, a, g
        xr, xw, xk, xv, xa, xg = (sx * mixes[j][:, None] + xx for j in range(6))
        r = self.mat(p + "att.receptance.weight").mul(xr)
        g = self.mat(p + "att.g2").mul(sigmoid(self.mat(p + "att.g1").mul(xg)))
        a = sigmoid(self.mat(p + "att.a2").mul(self.mat(p + "att.a1").mul(xa)) + self.vec(p + "att.a0")[:, None])
        w = self.mat(p + "att.w2").mul(np.tanh(self.mat(p + "att.w1").mul(xw)).astype(F32)) + self.vec(p + "att.w0")[:, None]
        w = np.exp(sigmoid(w) * F32(-0.606531), dtype=F32)
        k = self.mat(p + "att.key.weight").mul(xk)
        kk = (k * self.vec(p + "att.k_k")[:, None]).reshape(H, S, T)
        # rwkv_l2norm_impl (rwkv_operators.inc:40-82): x / max(||x||, 1e-12) per head
        nrm = np.sqrt((kk * kk).sum(axis=1, dtype=F32), dtype=F32)
        kk = (kk * (F32(1.0) / np.maximum(nrm, F32(1e-12)))[:, None, :]).reshape(C, T).astype(F32)
        ka = k * self.vec(p + "att.k_a")[:, None]
        k = k + (a * ka - ka)
        v = self.mat(p + "att.value.weight").mul(xv)
        if v_first is None:
            v_first = v
        else:
            gate = sigmoid(self.mat(p + "att.v2").mul(self.mat(p + "att.v1").mul(xv)) + self.vec(p + "att.v0")[:, None])
            v = v + (v_first - v) * gate
        y = self._wkv7(ls, r, w, k, v, -kk, kk * a)
        y = group_norm(y, H, 64e-5)
        y = y * self.vec(p + "att.ln_x.weight")[:, None] + self.vec(p + "att.ln_x.bias")[:, None]
        rk = (k * r * self.vec(p + "att.r_k")[:, None]).reshape(H, S, T).sum(axis=1, dtype=F32)  # ggml_sum_rows
        y = y + (v.reshape(H, S, T) * rk[:, None, :]).reshape(C, T)
        y = y * g
        return self.mat(p + "att.output.weight").mul(y), v_first

    def _ffn_v7(self, p, x, ls):
        """rwkv_ffn_v7 (rwkv_graph.inc:533-543)."""
        xx, prev = self._carry(p, "ln2", x, ls, 0)
        xk = (prev - xx) * self.vec(p + "ffn.x_k")[:, None] + xx
        k = np.square(np.maximum(self.mat(p + "ffn.key.weight").mul(xk), 0))
        return self.mat(p + "ffn.value.weight").mul(k)

    # chunked evaluation: rwkv_eval_sequence_in_chunks (rwkv_eval.inc:158-222)
    def eval_sequence_in_chunks(self, tokens, chunk_size, state_in=None, want_logits=True):
        tokens = list(tokens)
        state = self.init_state() if state_in is None else np.array(state_in, dtype=F32, copy=True)
        n_full, rem = divmod(len(tokens), chunk_size)
        logits, off = None, 0
        for c in range(n_full):
            last = c == n_full - 1 and rem == 0
            logits, state = self.eval_sequence(tokens[off:off + chunk_size], state, want_logits and last)
            off += chunk_size
        if rem:
            logits, state = self.eval_sequence(tokens[off:], state, want_logits)
        return logits, state


# --------------------------------------------------------------------------------------------
# Byte model: algorithmic bytes read per token at T=1
# --------------------------------------------------------------------------------------------

def bytes_per_token(path, with_logits=True):
    """Sum of rwkv_tensor_nbytes (rwkv_utilities.inc:1-3) over every tensor read for one token, one embedding
    row, the head + ln_out only when logits are requested, plus the recurrent state read + written (fp32)."""
    m = OracleModel(path)
    total = 0
    for name, t in m.file.tensors.items():
        if name == "emb.weight":
            total += len(t.raw) // t.ne3[1]
        elif name in ("head.weight", "ln_out.weight", "ln_out.bias"):
            total += len(t.raw) if with_logits else 0
        else:
            total += len(t.raw)
    return total + 2 * 4 * m.state_len
