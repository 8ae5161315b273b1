"""BERT-base encoder on the B200 kernels (forward + hand-scheduled backward).

Reference path: encoder.py:112-134 (TextEncoder -> transformers.BertModel.last_hidden_state),
mibf_net/bert.py:6-13 and ConNexT/models/BERT.py:13-21 (CLS row of the same).  transformers' BertModel is
only the parameter container (state_dict keys `embeddings.*`, `encoder.layer.N.*`, `pooler.*`).

Per layer: fused QKV GEMM (the three weight matrices are laid out back to back in the flat buffer),
fused masked-softmax attention, output GEMM with bias + dropout + residual in the epilogue, LayerNorm,
FFN1 GEMM with bias + GELU(erf) epilogue (pre-activation kept for the backward), FFN2 GEMM with bias +
dropout + residual, LayerNorm.  Activations bf16 [B*S, features]; LN statistics / parameters fp32.
The pooler is never evaluated (the reference discards it; its parameters get no gradient there either).
"""
import contextlib
import math

import torch

from . import ops
from . import runtime


def qkv_groups(bert):
    """Parameter groups that must be contiguous in the ParamStore (fused QKV operand)."""
    groups = []
    for layer in bert.encoder.layer:
        a = layer.attention.self
        groups.append([a.query.weight, a.key.weight, a.value.weight])
        groups.append([a.query.bias, a.key.bias, a.value.bias])
    return groups


class BertEngine:
    def __init__(self, store, bert):
        self.store, self.bert = store, bert
        cfg = bert.config
        self.H = cfg.num_attention_heads
        self.C = cfg.hidden_size
        self.D = self.C // self.H
        self.eps = cfg.layer_norm_eps
        self.p_hidden = float(cfg.hidden_dropout_prob)
        self.p_attn = float(cfg.attention_probs_dropout_prob)
        if cfg.hidden_act not in ("gelu",):
            raise ValueError(f"unsupported BERT activation {cfg.hidden_act!r} (erf GELU only)")
        self.step_seed = 0x5EED0000
        # called with the layer index right after that encoder layer's backward kernels have been enqueued (its parameter
        # gradients are final): the data-parallel trainer starts the layer group's gradient bucket early
        self.on_layer_backward_done = None
        self._s64 = None   # fp64 [2, K] column-sum workspace of the bias-gradient epilogue
        self.layers = []
        C = self.C
        for layer in bert.encoder.layer:
            a = layer.attention.self
            wqkv = store.fused([a.query.weight, a.key.weight, a.value.weight], (3 * C, C))
            bqkv = store.fused([a.query.bias, a.key.bias, a.value.bias], (3 * C,))
            self.layers.append(dict(
                wqkv=wqkv, bqkv=bqkv, qkv_params=[a.query.weight, a.key.weight, a.value.weight],
                wo=layer.attention.output.dense, ln1=layer.attention.output.LayerNorm,
                wi=layer.intermediate.dense, wo2=layer.output.dense, ln2=layer.output.LayerNorm))

    # ------------------------------------------------------------------ forward
    def forward(self, input_ids, attention_mask, training, need_grad, taps=()):
        """taps: 1-based encoder-layer numbers whose OUTPUT hidden state is returned as well (HF `hidden_states[n]`,
        `output_hidden_states=True`): forward returns (last_hidden, ctx, {n: hidden_n}) when taps is non-empty."""
        st = self.store
        B, S = input_ids.shape
        T, C, H, D = B * S, self.C, self.H, self.D
        emb = self.bert.embeddings
        ids = input_ids.reshape(-1).contiguous()
        mask8 = None
        if attention_mask is not None:
            mask8 = (attention_mask != 0).to(torch.uint8).contiguous()
        ph = self.p_hidden if training else 0.0
        pa = self.p_attn if training else 0.0
        if training:
            self.step_seed += 1000
        seed = self.step_seed
        e = ops.embed_gather(ids, None, emb.word_embeddings.weight.data, emb.position_embeddings.weight.data,
                             emb.token_type_embeddings.weight.data, S)
        x, _, e_mean, e_rstd = ops.layernorm_fwd(e, emb.LayerNorm.weight.data, emb.LayerNorm.bias.data, self.eps,
                                                 drop_p=ph, seed=seed + 1, save_stats=need_grad)
        ctx = dict(B=B, S=S, ids=ids, mask8=mask8, e=e, e_mean=e_mean, e_rstd=e_rstd, ph=ph, pa=pa, seed=seed,
                   layers=[]) if need_grad else None
        scale = 1.0 / math.sqrt(D)
        tapped = {}
        for li, L in enumerate(self.layers):
            s0 = seed + 10 * (li + 1)
            qkv = ops.gemm(x, L["wqkv"][1], bias=L["bqkv"][0])
            att, lse = ops.attention_fwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], B, H, S, S, D, scale,
                                         key_mask=mask8, drop_p=pa, seed=s0 + 1)
            pre1 = ops.gemm(att, st.w16(L["wo"].weight), bias=L["wo"].bias.data, residual=x, dropout_p=ph, dropout_seed=s0 + 2)
            h1, _, m1, r1 = ops.layernorm_fwd(pre1, L["ln1"].weight.data, L["ln1"].bias.data, self.eps, save_stats=need_grad)
            ipre = torch.empty((T, L["wi"].weight.shape[0]), device=x.device, dtype=torch.bfloat16) if need_grad else None
            # training: the epilogue also stores GELU'(pre-activation) (`ipre`), so the backward epilogue is one multiply
            inter = ops.gemm(h1, st.w16(L["wi"].weight), bias=L["wi"].bias.data,
                             act=ops.ACT_GELU_DERIV if need_grad else ops.ACT_GELU, aux_out=ipre)
            pre2 = ops.gemm(inter, st.w16(L["wo2"].weight), bias=L["wo2"].bias.data, residual=h1, dropout_p=ph,
                            dropout_seed=s0 + 3)
            h2, _, m2, r2 = ops.layernorm_fwd(pre2, L["ln2"].weight.data, L["ln2"].bias.data, self.eps, save_stats=need_grad)
            if need_grad:
                ctx["layers"].append(dict(x=x, qkv=qkv, att=att, lse=lse, pre1=pre1, m1=m1, r1=r1, h1=h1, ipre=ipre,
                                          inter=inter, pre2=pre2, m2=m2, r2=r2, s0=s0))
            x = h2
            if (li + 1) in taps:
                tapped[li + 1] = h2
        if taps:
            return x, ctx, tapped
        return x, ctx

    # ------------------------------------------------------------------ backward
    def _linear_bwd(self, dy, x_in, lin, need_dx=True, residual=None, aux_in=None, dact=ops.ACT_NONE, w16=None,
                    gw=None, gb=None, bias_done=False, colsum_to=None):
        """dy [T,N], x_in [T,K]: accumulates dW, db; returns dx = dy.W (* act'(aux_in)) (+ residual).
        colsum_to (fp32 [K] gradient view): += column sums of dx, taken in the dgrad GEMM's epilogue -- dx is the output
        gradient of the Linear below, so this is that Linear's bias gradient without a separate pass over dx."""
        st = self.store
        w = lin.weight if lin is not None else None
        w16 = st.w16(w) if w16 is None else w16
        N, K = w16.shape
        T = dy.shape[0]
        train = (w.requires_grad if w is not None else True)
        side = None
        if train:
            gw = st.g32(w) if gw is None else gw
            gb = (st.g32(lin.bias) if gb is None else gb)
            # weight / bias gradients do not feed the dgrad chain: a side stream lets them fill the dgrad kernel's last wave
            if runtime.OVERLAP_WGRAD and need_dx:
                side = runtime.fork_side()
            with torch.cuda.stream(side) if side is not None else contextlib.nullcontext():
                ops.gemm(dy, x_in, a_mn=True, b_mn=True, out=gw, accumulate=True, split_k=-1, M=N, N=K, K=T)
                if not bias_done:      # (the LayerNorm backward that produced dy already summed its columns)
                    ops.col_stats(dy, sum32=gb)
        if not need_dx:
            return None
        s64 = None
        if colsum_to is not None:
            if self._s64 is None or self._s64.shape[1] < K:
                self._s64 = torch.zeros((2, K), device=dy.device, dtype=torch.float64)   # kept zero by sum64_to_grad
            s64 = self._s64[:, :K]
        dx = ops.gemm(dy, w16, b_mn=True, residual=residual, aux_in=aux_in, dact=dact, M=T, N=K, K=N,
                      colsum=None if s64 is None else s64[0], colsumsq=None if s64 is None else s64[1])
        if s64 is not None:
            ops.sum64_to_grad(s64[0], s64[1], colsum_to)
        if side is not None:
            runtime.join_side(side)
        return dx

    def backward(self, ctx, dh, dtaps=None):
        """dh: grad of the last hidden state, [B*S, C] bf16 (None = zero).  dtaps: {layer number: grad of that layer's output
        hidden state} for the tapped hidden states of `forward(..., taps=...)`."""
        st = self.store
        B, S = ctx["B"], ctx["S"]
        C, H, D = self.C, self.H, self.D
        ph, pa = ctx["ph"], ctx["pa"]
        scale = 1.0 / math.sqrt(D)
        d = dh
        dtaps = dtaps or {}
        for li in range(len(self.layers) - 1, -1, -1):
            g_tap = dtaps.get(li + 1)
            if g_tap is not None:
                d = g_tap if d is None else ops.axpby_bf16(d, g_tap.contiguous(), 1.0, 1.0)
            if d is None:
                continue   # nothing downstream of this layer needs a gradient
            L, R = self.layers[li], ctx["layers"][li]
            s0 = R["s0"]
            ln2, ln1 = L["ln2"], L["ln1"]
            tr2 = ln2.weight.requires_grad
            b2 = st.g32(L["wo2"].bias) if L["wo2"].weight.requires_grad else None
            dpre2, dpre2_d, _ = ops.layernorm_bwd(d, R["pre2"], R["m2"], R["r2"], ln2.weight.data,
                                                  st.g32(ln2.weight) if tr2 else None, st.g32(ln2.bias) if tr2 else None,
                                                  drop2_p=ph, seed2=s0 + 3, want_dx_drop=ph > 0, dbias=b2)
            g2 = dpre2_d if ph > 0 else dpre2
            gbi = st.g32(L["wi"].bias) if L["wi"].weight.requires_grad else None
            dipre = self._linear_bwd(g2, R["inter"], L["wo2"], aux_in=R["ipre"], dact=ops.ACT_MUL, bias_done=True, colsum_to=gbi)
            dh1 = self._linear_bwd(dipre, R["h1"], L["wi"], residual=dpre2, bias_done=gbi is not None)
            tr1 = ln1.weight.requires_grad
            b1 = st.g32(L["wo"].bias) if L["wo"].weight.requires_grad else None
            dpre1, dpre1_d, _ = ops.layernorm_bwd(dh1, R["pre1"], R["m1"], R["r1"], ln1.weight.data,
                                                  st.g32(ln1.weight) if tr1 else None, st.g32(ln1.bias) if tr1 else None,
                                                  drop2_p=ph, seed2=s0 + 2, want_dx_drop=ph > 0, dbias=b1)
            g1 = dpre1_d if ph > 0 else dpre1
            datt = self._linear_bwd(g1, R["att"], L["wo"], bias_done=True)
            qkv = R["qkv"]
            dqkv = torch.empty_like(qkv)
            ops.attention_bwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], R["att"], datt, R["lse"], B, H, S, S, D, scale,
                              key_mask=ctx["mask8"], drop_p=pa, seed=s0 + 1,
                              dq=dqkv[:, :C], dk=dqkv[:, C:2 * C], dv=dqkv[:, 2 * C:])
            wq = L["qkv_params"][0]
            if wq.requires_grad:
                d = self._linear_bwd(dqkv, R["x"], None, residual=dpre1, w16=L["wqkv"][1], gw=L["wqkv"][2], gb=L["bqkv"][2])
            else:
                d = ops.gemm(dqkv, L["wqkv"][1], b_mn=True, residual=dpre1)
            if self.on_layer_backward_done is not None:
                self.on_layer_backward_done(li)
        emb = self.bert.embeddings
        if d is None or not emb.word_embeddings.weight.requires_grad:
            return
        _, _, de = ops.layernorm_bwd(d, ctx["e"], ctx["e_mean"], ctx["e_rstd"], emb.LayerNorm.weight.data,
                                     st.g32(emb.LayerNorm.weight), st.g32(emb.LayerNorm.bias), dx_bf16=False, dx_f32=True,
                                     drop_p=ph, seed=ctx["seed"] + 1)
        ops.embed_scatter(de, ctx["ids"], None, st.g32(emb.word_embeddings.weight), st.g32(emb.position_embeddings.weight),
                          st.g32(emb.token_type_embeddings.weight), S)
