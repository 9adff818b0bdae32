"""Host-side engine of the hot path: owns the flat parameter arena, the packed bf16 compute weights and
the activation workspaces, and sequences the libomlm_b200 kernels for forward, backward and the
optimiser step.  PyTorch supplies device memory and streams only; every arithmetic op is a call into
the C ABI (open_musiclm_b200.lib).

HBM layout
  arena_p / arena_g / adam_m / adam_v : one fp32 buffer each, parameters ordered
        [embeddings | logit heads | layer matrices | rel-pos MLP matrices]   <- weight-decayed (ndim >= 2)
        [start tokens | gammas, scales, MLP biases]                          <- not decayed
        every nn.Parameter of the module is a view into arena_p (and its .grad into arena_g).
  packed weights (16-bit, refreshed after every parameter update):
        wq [h*64, d], wkv [128, d], wo [d, h*64], w1 [2*Fp, d] (value rows | gate rows, zero padded),
        w2 [d, Fp], logit heads [q, Cp, d];  conv taps fp32 [2*Fp, 3], inner gamma fp32 [Fp].
  activations: residual stream fp32 [M, d]; GEMM operands 16-bit; attention statistics fp32.

16-bit operand formats (csrc/common.cuh): tcgen05 kind::f16 runs fp16 and bf16 at the same rate, but both operands
of one MMA must share the format (fp16 x bf16 faults on B200 -- measured).  The FORWARD GEMMs whose operands are bounded
by construction -- LayerNorm outputs (xn, xn2, hn, xf) against weights, and the FFN activations between them (u, h) --
run in fp16 (11-bit significand: 8x less operand rounding than bf16; this is what keeps the logits within 1e-2 of the
fp32 reference at 24 layers: 3.6e-3 instead of 1.25e-2).  Everything that touches an unbounded range stays bf16: every
BACKWARD GEMM (gradients), the K/V projection of the raw residual stream, attention and its output projection.  So
the weights are packed twice (fp16 for forward, bf16 for backward) and the saved LayerNorm outputs carry a bf16
duplicate for the weight-gradient GEMMs.  OMLM_ACT16=bf16 switches the whole path back to bf16 (diagnostics only).
"""
import math
import os
from typing import List, Optional

import torch

from . import lib

CE_IGNORE = -100


def _round_up(x, m):
    return (x + m - 1) // m * m


class _Plan:
    """Static shape bookkeeping for one (batch, sequence lengths) configuration."""

    def __init__(self, eng: "Engine", B: int, n_tok: List[int], device):
        self.B = B
        self.n_tok = list(n_tok)
        S = len(n_tok)
        self.N = sum(n + 1 for n in n_tok)
        self.M = B * self.N
        self.pos0 = list(itertools_accumulate([0] + [n + 1 for n in n_tok[:-1]]))
        # logits positions per sequence (open_musiclm.py:149-156)
        self.n_out = [n_tok[s] if s < S - 1 else n_tok[s] + 1 for s in range(S)]
        # head groups: (s, qi) -> rows ordered (b, t), position p = qi + q*t
        self.groups = []
        dest = torch.full((B, self.N), -1, dtype=torch.int32)
        base = 0
        self.seq_row_index = []   # per sequence: [B, n_out] -> row in the permuted buffers
        for s in range(S):
            q = eng.seqs[s].num_quantizers
            idx = torch.empty(B, self.n_out[s], dtype=torch.int64)
            for qi in range(min(q, self.n_out[s])):
                cnt = (self.n_out[s] - qi + q - 1) // q
                rows = base + torch.arange(B)[:, None] * cnt + torch.arange(cnt)[None, :]
                pos = qi + q * torch.arange(cnt)
                dest[:, self.pos0[s] + pos] = rows.to(torch.int32)
                idx[:, pos] = rows
                self.groups.append((s, qi, cnt, base))
                base += B * cnt
            self.seq_row_index.append(idx.to(device))
        self.rows_total = base
        self.dest_row = dest.reshape(-1).to(device)
        # absolute position embeddings (open_musiclm.py:134-136): token t of sequence s also gets row t of its
        # position table; start tokens get none.  Static per shape, so it is built here once.
        self.src_row2 = None
        if eng.abs_pos:
            r2 = torch.full((B, self.N), -1, dtype=torch.int32)
            for s in range(S):
                if n_tok[s] > eng.max_abs_pos:
                    raise IndexError(f"sequence {s} has {n_tok[s]} tokens but max_absolute_position_embeddings is {eng.max_abs_pos}")
                r2[:, self.pos0[s] + 1:self.pos0[s] + 1 + n_tok[s]] = eng.abs_row_base[s] + torch.arange(n_tok[s], dtype=torch.int32)
            self.src_row2 = r2.reshape(-1).to(device)


def itertools_accumulate(xs):
    t = 0
    out = []
    for x in xs:
        t += x
        out.append(t)
    return out


class Engine:
    def __init__(self, module):
        lib.load()
        self.m = module
        dev = module.device
        if dev.type != "cuda":
            raise lib.OmlmError("open_musiclm_b200 needs a CUDA (sm_100a) device: move the module with .to('cuda') "
                                "before calling it - there is no CPU fallback")
        lib.device_check()
        self.dev = dev
        self.seqs = module.token_sequences
        self.d, self.L, self.h = module.dim, module.depth, module.heads
        self.HD = self.h * 64
        self.use_conv_ff = bool(getattr(module, "use_conv_ff", True))
        self.bias_type = getattr(module, "relative_position_bias_type", "continuous")
        self.abs_pos = module.absolute_position_embeddings is not None
        self.max_abs_pos = int(getattr(module, "max_absolute_position_embeddings", 0))
        # feed-forward flavour (transformer.py:140-161): state_dict suffixes of (pre-norm gamma, up, conv, inner gamma, down).
        # The plain FeedForward runs through the same fused kernels with the conv taps pinned to (0, 0, 1).
        self.ffk = (dict(g1="2.0.gamma", w1="2.1.weight", conv="2.2.ds_conv.weight", gin="2.4.gamma", w2="2.6.weight") if self.use_conv_ff
                    else dict(g1="2.0.gamma", w1="2.1.weight", conv=None, gin="2.3.gamma", w2="2.5.weight"))
        self.F = int(self.d * 2 * 4 / 3) if self.use_conv_ff else int(self.d * 4)
        self.Fp = _round_up(self.F, 128)           # interleaved GEGLU layout: groups of 128 channels
        self.Hr = self.d // 2                      # rel-pos MLP width
        self.C = [s.codebook_size + 1 for s in self.seqs]
        self.Cp = [_round_up(c, 64) for c in self.C]
        self.drop_p = float(module.ff_dropout)
        self.alpha = float(module.grad_shrink_alpha)
        mode = os.environ.get("OMLM_ACT16", "fp16")
        if mode not in ("fp16", "bf16"):
            raise lib.OmlmError(f"OMLM_ACT16 must be 'fp16' or 'bf16', got {mode!r}")
        self.a16 = torch.float16 if mode == "fp16" else torch.bfloat16     # bounded forward operands (see module doc)
        self._build_arena()
        self._alloc_packed()
        self._packed_version = None
        self.bwd_max_ctas = 0      # > 0: CTA cap of the backward GEMMs (data parallel: leaves SMs to the overlapped NCCL kernels)
        self._plans = {}
        self._ws = {}
        self.seed = torch.zeros(1, dtype=torch.int64, device=dev)      # dropout / forgetful-mask seed (device resident)
        self.step_count = 0
        self.adam_m = None
        self.adam_v = None
        self.err_flag = torch.zeros(1, dtype=torch.int32, device=dev)  # latched by omlm_token_plan (token id out of range)
        self.loss_acc = torch.zeros(2, device=dev)
        self.sumsq = torch.zeros(1, device=dev, dtype=torch.float64)

    # ------------------------------------------------------------------------------------------ arena
    def _build_arena(self):
        named = [(n, p) for n, p in self.m.named_parameters()]
        is_row_table = lambda n: n.startswith("embeddings.") or n.startswith("absolute_position_embeddings.")
        emb = [(n, p) for n, p in named if is_row_table(n)]
        start = [(n, p) for n, p in named if n.startswith("start_tokens.")]
        decay = emb + [(n, p) for n, p in named if p.ndim >= 2 and not is_row_table(n)]
        nodecay = start + [(n, p) for n, p in named if p.ndim < 2 and not n.startswith("start_tokens.")]
        off, layout = 0, {}
        for n, p in decay:
            layout[n] = off
            off = _round_up(off + p.numel(), 64)
        self.n_decay = off
        emb_off = layout[emb[0][0]]
        # start tokens are addressed as rows of the embedding "table": keep them row-aligned to it
        off = emb_off + _round_up(off - emb_off, self.d)
        self.n_decay = off
        for n, p in nodecay:
            layout[n] = off
            off = _round_up(off + p.numel(), 64)
        self.n_params_arena = off
        self.layout = layout
        arena_p = torch.zeros(off, device=self.dev, dtype=torch.float32)
        arena_g = torch.zeros(off, device=self.dev, dtype=torch.float32)
        self.pview, self.gview = {}, {}
        for n, p in named:
            o = layout[n]
            v = arena_p[o:o + p.numel()].view(p.shape)
            v.copy_(p.data)
            p.data = v
            g = arena_g[o:o + p.numel()].view(p.shape)
            p.grad = g
            self.pview[n], self.gview[n] = v, g
        self.arena_p, self.arena_g = arena_p, arena_g
        # embedding table = [embeddings.0 | embeddings.1 | ... ] rows of d floats; start tokens further down
        self.emb_off = emb_off
        self.emb_row_base, self.abs_row_base = [], []
        for n, p in emb:
            assert (layout[n] - emb_off) % self.d == 0
            (self.emb_row_base if n.startswith("embeddings.") else self.abs_row_base).append((layout[n] - emb_off) // self.d)
        self.start_row = [(layout[n] - emb_off) // self.d for n, _ in start]
        self.table = arena_p[emb_off:]
        self.dtable_emb = arena_g[emb_off:]
        self._param_list = [p for _, p in named]

    def params_version(self):
        return sum(p._version for p in self._param_list)

    def _alloc_packed(self):
        dev, bf, a16 = self.dev, torch.bfloat16, self.a16
        d, HD, Fp = self.d, self.HD, self.Fp
        dual = a16 != bf        # forward operands in fp16, backward operands (suffix _b) in bf16; one buffer when equal
        self.pk = []
        for _ in range(self.L):
            pk = dict(
                wq=torch.empty(HD, d, device=dev, dtype=a16), w1=torch.empty(2 * Fp, d, device=dev, dtype=a16),
                w2=torch.empty(d, Fp, device=dev, dtype=a16),
                wkv_b=torch.empty(128, d, device=dev, dtype=bf), wo_b=torch.empty(d, HD, device=dev, dtype=bf),
                conv=torch.empty(2 * Fp, 3, device=dev), gin=torch.empty(Fp, device=dev))
            for k in ("wq", "w1", "w2"):
                pk[k + "_b"] = torch.empty_like(pk[k], dtype=bf) if dual else pk[k]
            self.pk.append(pk)
        if not self.use_conv_ff:
            for pk in self.pk:
                pk["conv"].zero_()
                pk["conv"][:, 2] = 1.0          # y[t] = u[t]: no depthwise conv in FeedForward (transformer.py:152-161)
        self.pk_logit = [torch.empty(s.num_quantizers, cp, d, device=dev, dtype=a16) for s, cp in zip(self.seqs, self.Cp)]
        self.pk_logit_b = [torch.empty_like(t, dtype=bf) if dual else t for t in self.pk_logit]
        self._pack_table = None
        self.pk_rp = [torch.empty(self.Hr, 3 * self.Hr, device=dev, dtype=torch.bfloat16) for _ in range(2)]   # rel-pos MLP layers 1, 2: [hi|lo|hi]

    def refresh_packed(self, force=False):
        ver = self.params_version()
        if not force and ver == self._packed_version:
            return
        d, HD, F, Fp = self.d, self.HD, self.F, self.Fp
        pv = self.pview
        if self._pack_table is None:      # the job table is built once: arena views and packed buffers never move
            tab = lib.PackTable(self.dev)
            for l, pk in enumerate(self.pk):
                p = f"transformer.layers.{l}."
                fk = self.ffk
                dual = pk["wq"] is not pk["wq_b"]       # fp16 forward copy + bf16 backward copy from ONE read of the matrix
                tab.add(pv[p + "0.to_q.weight"], d, HD, d, pk["wq"], HD, d, dst2=pk["wq_b"] if dual else None)
                tab.add(pv[p + fk["w1"]], d, 2 * F, d, pk["w1"], 2 * Fp, d, split_dst=-1, split_src=F, dst2=pk["w1_b"] if dual else None)
                tab.add(pv[p + fk["w2"]], F, d, F, pk["w2"], d, Fp, dst2=pk["w2_b"] if dual else None)
                tab.add(pv[p + "0.to_kv.weight"], d, 128, d, pk["wkv_b"], 128, d)
                tab.add(pv[p + "0.to_out.0.weight"], HD, d, HD, pk["wo_b"], d, HD)
                if fk["conv"] is not None:
                    tab.add(pv[p + fk["conv"]], 3, 2 * F, 3, pk["conv"], 2 * Fp, 3, split_dst=-1, split_src=F)
                tab.add(pv[p + fk["gin"]], F, 1, F, pk["gin"], 1, Fp)
            for s, seq in enumerate(self.seqs):
                # [q, C, d] -> [q, Cp, d]: every head padded with zero rows
                dual = self.pk_logit[s] is not self.pk_logit_b[s]
                tab.add(pv[f"logit_weights.{s}"], d, seq.num_quantizers * self.C[s], d, self.pk_logit[s].view(-1, d),
                        seq.num_quantizers * self.Cp[s], d, split_dst=self.Cp[s], split_src=self.C[s],
                        dst2=self.pk_logit_b[s].view(-1, d) if dual else None)
            self._pack_table = tab
        self._pack_table.run()
        if self.bias_type == "continuous":
            for j in (1, 2):
                lib.split3_bf16(pv[f"transformer.rel_pos_bias.net.{j}.0.weight"], self.pk_rp[j - 1], weight_mode=True)
        self._packed_version = ver

    def grad_bucket_plan(self, min_elems=4 << 20):
        """All-reduce buckets of the gradient arena in backward-completion order (dist_utils.plan_buckets)."""
        from .dist_utils import plan_buckets
        sizes = {n: p.numel() for n, p in self.m.named_parameters()}
        return plan_buckets(self.layout, sizes, self.n_params_arena, self.L, min_elems)

    def check_errors(self):
        """Raises if a token id outside an embedding table was seen since the last check (nn.Embedding's IndexError;
        asynchronous like the reference's device-side assert on CUDA: this call synchronises)."""
        bits = int(self.err_flag.item())
        if bits:
            self.err_flag.zero_()
            bad = [s for s in range(len(self.seqs)) if bits >> s & 1]
            raise lib.OmlmError(f"token id out of range for the embedding table of sequence(s) {bad} "
                                f"(valid ids: 0..codebook_size, or the pad id at quantizer-0 positions)")

    # ------------------------------------------------------------------------------------------ plans / workspaces
    _MAX_SHAPES = 8       # plans / workspaces kept (least recently used shapes are dropped: their HBM returns to torch)

    def plan(self, B, n_tok) -> _Plan:
        key = (B, tuple(n_tok))
        pl = self._plans.pop(key, None)
        if pl is None:
            pl = _Plan(self, B, n_tok, self.dev)
        self._plans[key] = pl                       # (re-)inserted last = most recently used
        while len(self._plans) > self._MAX_SHAPES:
            self._plans.pop(next(iter(self._plans)))
        return pl

    def workspace(self, pl: _Plan, train: bool):
        key = (pl.B, tuple(pl.n_tok), train)
        if key in self._ws:
            ws = self._ws.pop(key)
            self._ws[key] = ws                      # most recently used
            return ws
        while len(self._ws) >= self._MAX_SHAPES:
            self._ws.pop(next(iter(self._ws)))
        dev, bf, f32, a16 = self.dev, torch.bfloat16, torch.float32, self.a16
        M, d, HD, Fp, h = pl.M, self.d, self.HD, self.Fp, self.h
        E = lambda *shape, dt=bf: torch.empty(*shape, device=dev, dtype=dt)
        nl = self.L if train else 1
        ws = dict(
            x=[E(M, d, dt=f32) for _ in range(2 * nl + 1)],       # residual stream snapshots: x_l, x_mid_l, ..., x_L
            xn=[E(M, d) for _ in range(nl)], xraw=[E(M, d) for _ in range(nl)], st_a=[E(M, 2, dt=f32) for _ in range(nl)],
            q_raw=[E(M, HD) for _ in range(nl)], kv_raw=[E(M, 128) for _ in range(nl)],
            qn=[E(M, HD) for _ in range(nl)], kvn=[E(M, 128) for _ in range(nl)],
            o=[E(M, HD) for _ in range(nl)], lse=[E(M * h, dt=f32) for _ in range(nl)],
            xn2=[E(M, d) for _ in range(nl)], st_f=[E(M, 2, dt=f32) for _ in range(nl)],
            u=[E(M, 2 * Fp, dt=a16) for _ in range(nl)], hn=[E(M, Fp) for _ in range(nl)], st_i=[E(M, 2, dt=f32) for _ in range(nl)],
            keep=[E(M, Fp // 8, dt=torch.uint8) for _ in range(nl)],   # FFN dropout keep mask, 1 bit per element
            xf=E(max(pl.rows_total, 1), d), st_o=E(M, 2, dt=f32), h=E(M, Fp, dt=a16), rowsum=E(M, Fp // 128, 2, dt=f32),
            logits=[E(max(pl.B * c, 1), self.Cp[s], dt=f32) for (s, qi, c, b0) in pl.groups],
            # rel-pos MLP
            rp_in=E(pl.N, 1, dt=f32), rp_z=[E(pl.N, self.Hr, dt=f32) for _ in range(3)],
            rp_a=[E(pl.N, self.Hr, dt=f32) for _ in range(3)], table=E(h, pl.N, dt=f32),
            rp_a3=[E(pl.N, 3 * self.Hr) for _ in range(2)],
        )
        if a16 != bf:   # fp16 forward operands (transient: one buffer each); xn / xn2 / hn / xf above are then the bf16
            ws.update(xn16=E(M, d, dt=a16), xn2_16=E(M, d, dt=a16), hn16=E(M, Fp, dt=a16),   # duplicates kept for backward
                      xf16=E(max(pl.rows_total, 1), d, dt=a16))
        lib.arange_f32(ws["rp_in"])
        if self.bias_type == "none":
            ws["table"].zero_()                 # no bias is added (transformer.py:372-373): written once, never touched again
        elif self.bias_type == "t5":
            ws["ones"] = torch.ones(pl.N, device=dev, dtype=f32)
        if train:
            ws.update(
                dlogits=[E(max(pl.B * c, 1), self.Cp[s]) for (s, qi, c, b0) in pl.groups],
                dxf=E(max(pl.rows_total, 1), d), dx=[E(M, d, dt=f32) for _ in range(2)], dx_bf=E(M, d),
                dhn=E(M, Fp), rowstat=E(M, Fp // 128, 2, dt=f32), du=E(M, 2 * Fp), dxn=E(M, d), dxraw=E(M, d),
                d_o=E(M, HD), dqn=E(M, HD, dt=f32), dkvn=E(M, 128, dt=f32), dsum=E(M * h, dt=f32),
                dq_raw=E(M, HD), dkv_raw=E(M, 128), dtable=E(h, pl.N, dt=f32),
                rp_d0=E(pl.N, self.Hr, dt=f32), rp_d1=E(pl.N, self.Hr, dt=f32), rp_dz3=E(pl.N, 3 * self.Hr),
            )
        self._ws[key] = ws
        return ws

    # ------------------------------------------------------------------------------------------ forward
    # tile / split-K choice: minimise  waves x (k-blocks per unit x tile cost + epilogue)  over the 148 SMs
    _SMS = 148

    @classmethod
    def _tile_cost(cls, m, n, kb, bn, splits, epi):
        tiles = ((m + 127) // 128) * ((n + bn - 1) // bn)
        waves = (tiles * splits + cls._SMS - 1) // cls._SMS
        per_kb = 1.0 if bn == 128 else 2.0 / 1.2        # measured: 128x256 tiles are ~20 % more efficient per flop
        return waves * (((kb + splits - 1) // splits) * per_kb + epi * (bn / 128.0))

    @classmethod
    def _bn_for(cls, m, n, k):
        kb = (k + 63) // 64
        if n <= 128:
            return 128
        return min((128, 256), key=lambda bn: cls._tile_cost(m, n, kb, bn, 1, 6.0))

    @staticmethod
    def _bn(N):
        return 256 if N % 256 == 0 or N >= 2048 else 128

    def build_bias_table(self, ws, N):
        """table[h, delta] for delta = i - j in [0, N) of the configured relative position bias (transformer.py:366-373)."""
        if self.bias_type == "continuous":
            self._relpos_table(ws, N)
        elif self.bias_type == "t5":
            # T5RelativePositionBias (transformer.py:69-117) is fed i - j and negates it, so every causally visible pair
            # falls into bucket 0: the table is the constant row 0 of the bucket embedding, one value per head
            w = self.pview["transformer.rel_pos_bias.relative_attention_bias.weight"]       # [32, h]
            lib.sgemm_small(w, (1, 1), ws["ones"], (1, 1), ws["table"], (ws["table"].stride(0), 1), self.h, N, 1)
        # 'none': the table stays zero

    def bias_table_backward(self, ws, N):
        if self.bias_type == "continuous":
            self._relpos_backward(ws, N)
        elif self.bias_type == "t5":
            gw = self.gview["transformer.rel_pos_bias.relative_attention_bias.weight"]      # bucket 0 collects every delta
            lib.colsum(ws["dtable"], 1, ws["dtable"].stride(0), gw[0], N, self.h, accumulate=True)

    def _relpos_table(self, ws, N):
        """RelativePositionBias MLP on the causal distances 0..N-1 -> table[h, N] (transformer.py:55-67).
        The two Hr x Hr layers run on the tcgen05 GEMM with bf16x3-split operands (fp32-class accuracy: the
        table reaches |b| ~ 100 and dominates the logits); the rank-1 first layer and the h-wide last layer are SIMT."""
        pv, Hr, h = self.pview, self.Hr, self.h
        pre = "transformer.rel_pos_bias.net."
        lib.sgemm_small(ws["rp_in"], (1, 1), pv[pre + "0.0.weight"], (1, 1), ws["rp_a"][0], (Hr, 1), N, Hr, 1,
                        Z=ws["rp_z"][0], bias=pv[pre + "0.0.bias"], act=1)
        for j in (1, 2):
            lib.split3_bf16(ws["rp_a"][j - 1], ws["rp_a3"][j - 1])
            lib.gemm(ws["rp_a3"][j - 1], self.pk_rp[j - 1], ws["rp_z"][j], block_n=128)
            lib.bias_silu(ws["rp_z"][j], pv[f"{pre}{j}.0.bias"], ws["rp_a"][j])
        # no split-K here: atomics would make the table, and through bf16 rounding every logit, depend on CTA timing
        lib.sgemm_small(ws["rp_a"][2], (Hr, 1), pv[pre + "3.weight"], (1, Hr), ws["table"], (1, N), N, h, Hr, bias=pv[pre + "3.bias"])

    def forward_core(self, pl: _Plan, ws, src_row, key_mask, train: bool, groups_wanted=None, drop: bool = False, capture=None):
        """Runs embeddings -> depth x (attention, conv-FFN) -> final LN -> logit heads.  Activations stay in `ws`.
        capture (decode.py): receives each layer's K/V rows and pre-conv FFN rows (the generation caches of the prompt)."""
        self.refresh_packed()
        B, N, M, d, h, HD, F, Fp = pl.B, pl.N, pl.M, self.d, self.h, self.HD, self.F, self.Fp
        pv = self.pview
        x = ws["x"]
        lib.embed_gather(self.table, src_row, x[0], pl.src_row2)
        self.build_bias_table(ws, N)
        drop_p = self.drop_p if drop else 0.0
        f16 = self.a16 != torch.bfloat16
        dup = f16 and train            # the backward pass needs bf16 duplicates of the fp16 forward operands
        for l in range(self.L):
            i = l if train else 0
            xa, xm, xo = (x[2 * l], x[2 * l + 1], x[2 * l + 2]) if train else (x[0], x[1], x[0])
            p, pk = f"transformer.layers.{l}.", self.pk[l]
            xn = ws["xn16"] if f16 else ws["xn"][i]
            lib.layernorm_fwd(xa, pv[p + "0.norm.gamma"], xn, ws["xraw"][i], ws["st_a"][i], ycopy=ws["xn"][i] if dup else None)
            lib.gemm(xn, pk["wq"], ws["q_raw"][i], block_n=self._bn_for(M, HD, d))
            lib.gemm(ws["xraw"][i], pk["wkv_b"], ws["kv_raw"][i], block_n=128)
            lib.qk_l2norm_fwd(ws["q_raw"][i], ws["kv_raw"][i], pv[p + "0.q_scale"], pv[p + "0.k_scale"], ws["qn"][i], ws["kvn"][i], h)
            if capture is not None:
                capture.after_kv(l, ws["kvn"][i])
            lib.attn_fwd_tc(ws["qn"][i], ws["kvn"][i], ws["table"], key_mask, ws["o"][i], ws["lse"][i], B, N, h)
            lib.gemm(ws["o"][i], pk["wo_b"], xm, addend=xa, block_n=self._bn_for(M, d, HD))
            xn2 = ws["xn2_16"] if f16 else ws["xn2"][i]
            lib.layernorm_fwd(xm, pv[p + self.ffk["g1"]], xn2, None, ws["st_f"][i], ycopy=ws["xn2"][i] if dup else None)
            lib.gemm_ffn_up(xn2, pk["w1"], pk["conv"], ws["u"][i], ws["h"], ws["rowsum"], N, Fp)   # conv + GEGLU in the epilogue
            if capture is not None:
                capture.after_u(l, ws["u"][i])
            hn = ws["hn16"] if f16 else ws["hn"][i]
            lib.ffn_norm_fwd(ws["h"], ws["rowsum"], pk["gin"], hn, ws["st_i"][i], F, Fp, drop_p, self.seed, l,
                             keep_bits=ws["keep"][i] if drop_p > 0 else None, hn_copy=ws["hn"][i] if dup else None)
            lib.gemm(hn, pk["w2"], xo, addend=xm, block_n=self._bn_for(M, d, Fp))
        x_last = x[2 * self.L] if train else x[0]
        xf = ws["xf16"] if f16 else ws["xf"]
        lib.layernorm_fwd(x_last, pv["transformer.norm.gamma"], xf, None, ws["st_o"], pl.dest_row, ycopy=ws["xf"] if dup else None)
        for gi, (s, qi, cnt, base) in enumerate(pl.groups):
            if groups_wanted is not None and s not in groups_wanted:
                continue
            rows = B * cnt
            lib.gemm(xf[base:base + rows], self.pk_logit[s][qi], ws["logits"][gi], block_n=128)

    # ------------------------------------------------------------------------------------------ backward
    def _wgrad(self, dy, x, gout, m, n, **kw):
        """gout[m, n] += dy[rows, m]^T x[rows, n]   (both operands MN-major, fp32 accumulate into the grad arena)."""
        k = dy.shape[0]
        kb = (k + 63) // 64
        best = None
        for bn in (128, 256):
            for s in range(1, max(1, min(kb // 8, 48)) + 1):
                c = self._tile_cost(m, n, kb, bn, s, 14.0 if s > 1 else 10.0)
                if best is None or c < best[0]:
                    best = (c, bn, s)
        _, bn, s = best
        if s > 1:
            lib.gemm(dy, x, gout, a_mn=True, b_mn=True, M=m, N=n, K=k, splits=s, block_n=bn, max_ctas=self.bwd_max_ctas, **kw)
        else:
            lib.gemm(dy, x, gout, a_mn=True, b_mn=True, M=m, N=n, K=k, addend=gout, block_n=bn, max_ctas=self.bwd_max_ctas, **kw)

    def backward_core(self, pl: _Plan, ws, src_row, key_mask, groups_with_grad, drop: bool = False, on_ready=None):
        """Consumes ws['dlogits'] (bf16, permuted rows) and accumulates every parameter gradient into arena_g.
        on_ready(trigger): called when a group of gradients is final -- 'heads', 'layer<l>' (matrices of layer l), 'tail'
        (everything else) -- so that a data-parallel caller can start reducing it underneath the rest of the pass."""
        ready = on_ready if on_ready is not None else (lambda trigger: None)
        B, N, M, d, h, HD, F, Fp = pl.B, pl.N, pl.M, self.d, self.h, self.HD, self.F, self.Fp
        pv, gv = self.pview, self.gview
        x = ws["x"]
        drop_p = self.drop_p if drop else 0.0
        # ---- logit heads
        gset = frozenset(groups_with_grad)
        if ws.get("dxf_groups") != gset:      # rows of head groups without a gradient stay zero; the others are overwritten
            ws["dxf"].zero_()
            ws["dxf_groups"] = gset
        for gi, (s, qi, cnt, base) in enumerate(pl.groups):
            if s not in groups_with_grad:
                continue
            rows = B * cnt
            dl = ws["dlogits"][gi]
            lib.gemm(dl, self.pk_logit_b[s][qi], ws["dxf"][base:base + rows], b_mn=True, M=rows, N=d, K=self.Cp[s], block_n=128, max_ctas=self.bwd_max_ctas)
            self._wgrad(dl, ws["xf"][base:base + rows], gv[f"logit_weights.{s}"][qi], self.Cp[s], d, row_split=self.Cp[s], row_valid=self.C[s])
        ready("heads")
        dxa, dxb = ws["dx"]
        lib.layernorm_bwd(ws["dxf"], x[2 * self.L], ws["st_o"], pv["transformer.norm.gamma"], dxa, gv["transformer.norm.gamma"],
                          src_row=pl.dest_row, dx_bf16=ws["dx_bf"])
        ws["dtable"].zero_()
        for l in reversed(range(self.L)):
            p, pk = f"transformer.layers.{l}.", self.pk[l]
            xa, xm = x[2 * l], x[2 * l + 1]
            # ---- conv feed-forward
            # d_hn = dx W2 with the LayerNorm-backward row sums (against the saved hn) taken in the GEMM's epilogue
            fk = self.ffk
            keep = ws["keep"][l] if drop_p > 0 else None
            if Fp % 256 == 0:
                lib.gemm_rowstat(ws["dx_bf"], pk["w2_b"], ws["dhn"], ws["hn"][l], pk["gin"], ws["rowstat"], b_mn=True, M=M, N=Fp, K=d,
                                 keep_bits=keep, keep_scale=1.0 / (1.0 - drop_p) if drop_p > 0 else 1.0, max_ctas=self.bwd_max_ctas)
                parts = Fp // 128
            else:
                lib.gemm(ws["dx_bf"], pk["w2_b"], ws["dhn"], b_mn=True, M=M, N=Fp, K=d, block_n=self._bn_for(M, Fp, d), max_ctas=self.bwd_max_ctas)
                parts = 0
            # (the tile kernel runs right behind the GEMM that wrote dhn, while dhn is still in L2; the weight gradient after it)
            lib.ffn_mid_bwd(ws["dhn"], ws["hn"][l], ws["u"][l], ws["st_i"][l], pk["conv"], pk["gin"], ws["rowstat"], ws["du"],
                            gv[p + fk["gin"]], gv[p + fk["conv"]] if fk["conv"] is not None else None, B, N, F, Fp, drop_p,
                            keep_bits=keep, rowstat_parts=parts)
            self._wgrad(ws["dx_bf"], ws["hn"][l], gv[p + fk["w2"]], d, Fp, n_valid=F)
            lib.gemm(ws["du"], pk["w1_b"], ws["dxn"], b_mn=True, M=M, N=d, K=2 * Fp, block_n=self._bn_for(M, d, 2 * Fp), max_ctas=self.bwd_max_ctas)
            self._wgrad(ws["du"], ws["xn2"][l], gv[p + fk["w1"]], 2 * Fp, d, row_split=-1, row_valid=F)
            lib.layernorm_bwd(ws["dxn"], xm, ws["st_f"][l], pv[p + fk["g1"]], dxb, gv[p + fk["g1"]], dres=dxa, dx_bf16=ws["dx_bf"])
            # ---- attention
            lib.gemm(ws["dx_bf"], pk["wo_b"], ws["d_o"], b_mn=True, M=M, N=HD, K=d, block_n=self._bn_for(M, HD, d), max_ctas=self.bwd_max_ctas)
            self._wgrad(ws["dx_bf"], ws["o"][l], gv[p + "0.to_out.0.weight"], d, HD)
            lib.attn_bwd_tc(ws["qn"][l], ws["kvn"][l], ws["d_o"], ws["o"][l], ws["lse"][l], ws["table"], key_mask, ws["dsum"],
                            ws["dqn"], ws["dkvn"], ws["dtable"], B, N, h)
            lib.qk_l2norm_bwd(ws["dqn"], ws["dkvn"], ws["q_raw"][l], ws["kv_raw"][l], pv[p + "0.q_scale"], pv[p + "0.k_scale"],
                              ws["dq_raw"], ws["dkv_raw"], gv[p + "0.q_scale"], gv[p + "0.k_scale"], h)
            lib.gemm(ws["dq_raw"], pk["wq_b"], ws["dxn"], b_mn=True, M=M, N=d, K=HD, block_n=self._bn_for(M, d, HD), max_ctas=self.bwd_max_ctas)
            lib.gemm(ws["dkv_raw"], pk["wkv_b"], ws["dxraw"], b_mn=True, M=M, N=d, K=128, block_n=self._bn_for(M, d, 128), max_ctas=self.bwd_max_ctas)
            self._wgrad(ws["dq_raw"], ws["xn"][l], gv[p + "0.to_q.weight"], HD, d)
            self._wgrad(ws["dkv_raw"], ws["xraw"][l], gv[p + "0.to_kv.weight"], 128, d)
            ready(f"layer{l}")
            lib.layernorm_bwd(ws["dxn"], xa, ws["st_a"][l], pv[p + "0.norm.gamma"], dxa, gv[p + "0.norm.gamma"], dres=dxb, draw=ws["dxraw"],
                              dx_bf16=ws["dx_bf"])
        # ---- embeddings + start tokens (grad_shrink: utils.py:60-61)
        lib.embed_scatter_add(self.dtable_emb, src_row, dxa, self.alpha)
        if pl.src_row2 is not None:
            lib.embed_scatter_add(self.dtable_emb, pl.src_row2, dxa, self.alpha)
        self.bias_table_backward(ws, N)
        ready("tail")

    def _relpos_backward(self, ws, N):
        pv, gv, Hr, h = self.pview, self.gview, self.Hr, self.h
        pre = "transformer.rel_pos_bias.net."
        dT = ws["dtable"]                                   # [h, N]: dY[n, hh] = dT[hh, n]
        a3 = ws["rp_a"][2]
        lib.sgemm_small(dT, (N, 1), a3, (Hr, 1), gv[pre + "3.weight"], (Hr, 1), h, Hr, N, accumulate=True)      # dW4 = dY^T a3
        lib.colsum(dT, 1, N, gv[pre + "3.bias"], N, h, accumulate=True)
        d_cur, d_nxt = ws["rp_d0"], ws["rp_d1"]
        lib.sgemm_small(dT, (1, N), pv[pre + "3.weight"], (Hr, 1), d_cur, (Hr, 1), N, Hr, h)                      # da3 = dY W4
        for j in (2, 1, 0):
            lib.silu_bwd(d_cur, ws["rp_z"][j], d_cur)                                                              # dz_j (fp32, in place)
            lib.colsum(d_cur, Hr, 1, gv[f"{pre}{j}.0.bias"], N, Hr, accumulate=True)
            if j > 0:
                # bf16x3 products, as in the forward pass (x y ~ x_hi y_hi + x_hi y_lo + x_lo y_hi: fp32-class): these
                # gradients are sums of cancelling terms, so a plain bf16 operand rounding shows up amplified
                lib.split3_bf16(d_cur, ws["rp_dz3"])                                                               # [hi | hi | lo]
                dz_hi, dz_lo = ws["rp_dz3"][:, :Hr], ws["rp_dz3"][:, 2 * Hr:]
                a_hi, a_lo = ws["rp_a3"][j - 1][:, :Hr], ws["rp_a3"][j - 1][:, 2 * Hr:]                            # forward split of a_{j-1}
                w_hi, w_lo = self.pk_rp[j - 1][:, :Hr], self.pk_rp[j - 1][:, Hr:2 * Hr]                            # [hi | lo | hi]
                gw = gv[f"{pre}{j}.0.weight"]
                for dz, a in ((dz_hi, a_hi), (dz_hi, a_lo), (dz_lo, a_hi)):                                         # dW_j += dz^T a
                    lib.gemm(dz, a, gw, a_mn=True, b_mn=True, M=Hr, N=Hr, K=N, addend=gw, block_n=128)
                lib.gemm(dz_hi, w_hi, d_nxt, b_mn=True, M=N, N=Hr, K=Hr, block_n=128)                               # da = dz W
                lib.gemm(dz_hi, w_lo, d_nxt, b_mn=True, M=N, N=Hr, K=Hr, addend=d_nxt, block_n=128)
                lib.gemm(dz_lo, w_hi, d_nxt, b_mn=True, M=N, N=Hr, K=Hr, addend=d_nxt, block_n=128)
                d_cur, d_nxt = d_nxt, d_cur
            else:
                lib.sgemm_small(d_cur, (1, Hr), ws["rp_in"], (1, 1), gv[f"{pre}0.0.weight"], (1, 1), Hr, 1, N, accumulate=True)

    # ------------------------------------------------------------------------------------------ reference-API path
    def api_forward(self, all_token_ids, self_attn_mask, only_final):
        ids = [t.reshape(t.shape[0], -1).to(self.dev, torch.int64).contiguous() for t in all_token_ids]
        assert len(ids) == len(self.seqs)
        B = ids[0].shape[0]
        mask_in = None
        if self_attn_mask is not None:
            mask_in = self_attn_mask.to(self.dev).to(torch.uint8).contiguous()
        _, src_row, key_mask, _, n_tok = lib.token_plan(
            ids, [s.codebook_size for s in self.seqs], [s.num_quantizers for s in self.seqs], self.emb_row_base,
            self.start_row, append_eos=False, drop_last=False, mask_cond=False, mask_in=mask_in, want_labels=False,
            err_flag=self.err_flag)
        pl = self.plan(B, n_tok)
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self._param_list)
        wanted = {len(self.seqs) - 1} if only_final else set(range(len(self.seqs)))
        drop = self.m.training and self.drop_p > 0
        if drop:
            self.seed += 1
        outs = _ApiFunction.apply(self, pl, src_row, key_mask, need_grad, wanted, drop, *self._param_list)
        res, k = [], 0
        for s in range(len(self.seqs)):
            if s in wanted:
                res.append(outs[k]); k += 1
            else:
                res.append(None)
        return res

    def gather_logits(self, pl, ws, s):
        """Permuted group buffers -> [B, n_out, C] fp32 (re-layout only)."""
        parts = [ws["logits"][gi][:, :self.C[s]] for gi, g in enumerate(pl.groups) if g[0] == s]
        allrows = torch.cat(parts, 0)
        first = min(g[3] for g in pl.groups if g[0] == s)
        return allrows[(pl.seq_row_index[s] - first).reshape(-1)].view(pl.B, pl.n_out[s], self.C[s])

    def scatter_dlogits(self, pl, ws, s, grad):
        """[B, n_out, C] fp32 gradient -> permuted bf16 dlogits buffers (re-layout + cast only)."""
        first = min(g[3] for g in pl.groups if g[0] == s)
        total = sum(pl.B * g[2] for g in pl.groups if g[0] == s)
        buf = torch.zeros(total, self.Cp[s], device=self.dev, dtype=torch.bfloat16)
        buf[(pl.seq_row_index[s] - first).reshape(-1), :self.C[s]] = grad.reshape(-1, self.C[s]).to(torch.bfloat16)
        off = 0
        for gi, g in enumerate(pl.groups):
            if g[0] == s:
                n = pl.B * g[2]
                ws["dlogits"][gi].copy_(buf[off:off + n]); off += n


class _ApiFunction(torch.autograd.Function):
    """One autograd node for the whole TokenConditionedTransformer.forward: libomlm_b200 forward in
    forward(), libomlm_b200 backward in backward(); gradients are returned per parameter."""

    @staticmethod
    def forward(ctx, eng: Engine, pl, src_row, key_mask, need_grad, wanted, drop, *params):
        ws = eng.workspace(pl, need_grad)
        eng.forward_core(pl, ws, src_row, key_mask, need_grad, wanted, drop)
        ctx.eng, ctx.pl, ctx.src_row, ctx.key_mask, ctx.wanted, ctx.drop = eng, pl, src_row, key_mask, sorted(wanted), drop
        ctx.need_grad = need_grad
        # the saved activations live in the shape's workspace, not in the graph: a second forward of the same shape before
        # this call's backward would overwrite them -- remember which forward owns the workspace and check in backward
        ws["generation"] = ws.get("generation", 0) + 1
        ctx.ws, ctx.generation = ws, ws["generation"]
        outs = tuple(eng.gather_logits(pl, ws, s) for s in sorted(wanted))
        return outs

    @staticmethod
    def backward(ctx, *grads):
        eng, pl = ctx.eng, ctx.pl
        if not ctx.need_grad:
            raise RuntimeError("forward ran without gradient bookkeeping")
        ws = ctx.ws
        if ws.get("generation") != ctx.generation:
            raise RuntimeError("open_musiclm_b200: the activations of this forward pass were overwritten by a later forward of the "
                               "same shape (one workspace per shape): call backward() before the next forward, or use "
                               "HotPathTrainer for gradient accumulation")
        with_grad = set()
        for s, g in zip(ctx.wanted, grads):
            if g is not None:
                eng.scatter_dlogits(pl, ws, s, g)
                with_grad.add(s)
        # gradients are produced in a scratch copy of the arena so that autograd can accumulate them itself
        saved = eng.arena_g.clone()
        eng.arena_g.zero_()
        eng.backward_core(pl, ws, ctx.src_row, ctx.key_mask, with_grad, ctx.drop)
        fresh = eng.arena_g.clone()
        eng.arena_g.copy_(saved)
        outs = []
        for n, p in eng.m.named_parameters():
            o = eng.layout[n]
            outs.append(fresh[o:o + p.numel()].view(p.shape))
        return (None, None, None, None, None, None, None, *outs)
