"""Eval-flow composition of the hot path: the body of condGANTrainer.gen_exampleSRHL
(trainer_objective.py:134-146) -- text encoder -> mask -> G_SR_NET_low -> NetG_highweight --
as one callable.  This is the user-facing call bench.py times end to end."""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from . import nets, nets16
from .config import cfg
from .text_encoder import RNN_ENCODER


# TGSR_HEAD_STREAM=1: G_SR_NET_low's image heads on their own stream (measured: no gain, profiles/r2_experiments.md)
_HEAD_STREAM = int(__import__('os').environ.get('TGSR_HEAD_STREAM', '0'))


class SRPipeline:
    def __init__(self, n_words: int, device='cuda', variant: str = 'x8', low: str = 'lr', seed: Optional[int] = None,
                 randomize_bn: bool = False):
        """variant 'x8' -> model.py nets (TREE.BRANCH_NUM == 4), 'x16' -> models16.py
        (trainer_objective.py:74-88).  Weights are random-init (seeded) unless load_* is called."""
        if seed is not None:
            torch.manual_seed(seed)
        mod = nets if variant == 'x8' else nets16
        self.variant = variant
        self.device = torch.device(device)
        self.text_encoder = RNN_ENCODER(n_words, nhidden=cfg.TEXT.EMBEDDING_DIM)
        self.netGL = mod.G_SR_NET_low()
        self.netGH = mod.NetG_highweight(weightmap=False, low=low)
        if randomize_bn:
            randomize_bn_stats([self.netGL, self.netGH], seed or 0)
        for m in (self.text_encoder, self.netGL, self.netGH):
            m.to(self.device).eval()
        self._pinned = {}
        self._side = None
        self._copy_stream = None
        self._h2d_stream = None
        self._slots, self._tickets, self._ticket = {}, {}, 0

    def load_generators(self, sd_netG, sd_netGH):
        self.netGL.load_state_dict(sd_netG, strict=True)
        self.netGH.load_state_dict(sd_netGH, strict=True)
        self._graphs = {}                      # captured graphs bake the packed / folded weights of the old parameters

    @torch.no_grad()
    def forward_device(self, LR: torch.Tensor, captions: torch.Tensor, cap_lens, LRb: Optional[torch.Tensor] = None):
        """Inputs on the device.  Returns (fine_imgs list, fake_imgL list, attention_maps list)."""
        B = captions.shape[0]
        hidden = None                                                     # zero state (init_hidden), built in-kernel
        words_embs, sent_emb = self.text_encoder(captions, cap_lens, hidden)
        mask = (captions == 0)                                            # trainer_objective.py:136
        num_words = words_embs.size(2)
        if mask.size(1) > num_words:
            mask = mask[:, :num_words]
        fake_imgL, attention_maps, mu, logvar = self.netGL(LR, sent_emb, words_embs, mask)
        fine_im, a, one = self.netGH(LR, fake_imgL, LR if LRb is None else LRb)
        return fine_im, fake_imgL, attention_maps

    @torch.no_grad()
    def text_and_generate(self, LR, captions, cap_lens, tmax: int):
        """forward_device with max(cap_lens) known on the host: no device->host read anywhere.
        In the 16-bit modes the text-independent head of NetG_highweight (convin, 6 ResBlocks, upscale2x at 32x32 --
        small grids) runs on a second stream next to the text encoder (a latency-bound cluster kernel) and the
        first G_SR_NET_low stage; the streams join before the first high-frequency image head."""
        from . import fast, fmt_for, is_16bit
        gh, gl = self.netGH, self.netGL
        overlap = (is_16bit() and gh.low == 'lr' and LR.shape[2] % 16 == 0 and LR.shape[3] % 8 == 0
                   and gh.convin[0].weight.shape[0] == 64 and fast.stage_supported(gl.h_net1, LR.shape[2], LR.shape[3]))
        if not overlap:
            words_embs, sent_emb = self.text_encoder._forward_tmax(captions, cap_lens, tmax)
            mask = (captions == 0)[:, :tmax]
            fake_imgL, attention_maps, mu, logvar = gl(LR, sent_emb, words_embs, mask)
            fine_im, a, one = gh(LR, fake_imgL, LR)
            return fine_im, fake_imgL, attention_maps
        # Three independent chains start a step: the text encoder (a latency-bound cluster kernel; everything word-related
        # hangs off it), im2f of the first stage, and the whole 32x32 trunk of NetG_highweight.  Only the first is on the
        # critical path, so the other two run on their own streams and join where their results are consumed.
        cur = torch.cuda.current_stream(self.device)
        if self._side is None:
            self._side = (torch.cuda.Stream(self.device), torch.cuda.Stream(self.device), torch.cuda.Stream(self.device))
        s_trunk, s_im2f, s_head = self._side
        if not _HEAD_STREAM:
            s_head = None
        s_trunk.wait_stream(cur); s_im2f.wait_stream(cur)                    # fork before anything of this step is enqueued
        words_embs, sent_emb = self.text_encoder._forward_tmax(captions, cap_lens, tmax)   # critical chain first
        with torch.cuda.stream(s_im2f):
            mask = (captions == 0)[:, :tmax].to(torch.uint8).contiguous()     # trainer_objective.py:136-140
            h0 = fast.stage_im2f(gl.h_net1, LR, fmt_for('low'))
            ev_h0 = torch.cuda.Event()
            ev_h0.record(s_im2f)
        with torch.cuda.stream(s_trunk):
            h_high = fast.netg_high_trunk(gh, LR, None, fmt_for('high'))
        if self.variant == 'x8':                                         # model.py:34-78 / :264-298
            stages, heads = [gl.h_net1, gl.h_net2, gl.h_net3], [gl.img_net1, gl.img_net2, gl.img_net3]
            tail, a_val = [(gh.residual24, gh.upscale4x), (gh.residual48, gh.upscale8x)], 0.5
        else:                                                            # models16.py:5-39 / :145-179 (last stage reuses 48 / 8x)
            stages, heads = [gl.h_net1, gl.h_net2, gl.h_net3, gl.h_net4], [gl.img_net1, gl.img_net2, gl.img_net3, gl.img_net4]
            tail = [(gh.residual24, gh.upscale4x), (gh.residual48, gh.upscale8x), (gh.residual48, gh.upscale8x)]
            a_val = nets16._cached_float(gh, gh.a, '_tgsr_a')
        for st in stages:
            st.att.applyMask(mask)                                       # keep the modules' stateful contract

        def ca():                                                        # mu / logvar are only returned: off the critical path
            s_im2f.wait_stream(cur)
            with torch.cuda.stream(s_im2f):
                out = gl.ca_net(sent_emb)
            return out
        cur.wait_event(ev_h0)
        h0.buf.record_stream(cur); mask.record_stream(cur)
        # G_SR_NET_low's own image heads feed nothing inside that net: they run on a fourth stream and overlap the next stage;
        # NetG_high's heads wait for image k right where they add it (one*y + a*SRb[k])
        head_events: list = []
        fake_imgL, attention_maps, mu, logvar = fast.g_sr_net_low(gl, LR, sent_emb, words_embs, mask, stages, heads, False,
                                                                  fmt_for('low'), h0=h0, ca=ca, head_stream=s_head,
                                                                  head_events=head_events)
        cur.wait_stream(s_im2f)
        mu.record_stream(cur); logvar.record_stream(cur)
        cur.wait_stream(s_trunk)
        h_high.buf.record_stream(cur)
        fine_im = fast.netg_high_finish(gh, h_high, fake_imgL, a_val, 1.0, tail, srb_events=head_events or None)
        if s_head is not None:
            cur.wait_stream(s_head)                                      # join (a no-op after the last event wait)
        return fine_im, fake_imgL, attention_maps

    # ------------------------------------------------------------------ CUDA-graph replay
    @torch.no_grad()
    def forward_graph(self, LR, captions, cap_lens, tmax: int):
        """Same result as text_and_generate, but the ~50 kernel launches of one forward are captured
        once per input signature into a CUDA graph and replayed with a single launch.  Inputs are copied
        into the graph's static buffers unless they already ARE those buffers (see static_inputs()).
        The returned tensors are the graph's static outputs: they are overwritten by the next replay."""
        key = self._graph_key(tuple(LR.shape), tuple(captions.shape), tmax)
        e = self._graphs.get(key) if hasattr(self, '_graphs') else None
        if e is None:
            e = self._capture(key, LR, captions, cap_lens, tmax)
        if LR.data_ptr() != e['LR'].data_ptr():
            e['LR'].copy_(LR, non_blocking=True)
        if captions.data_ptr() != e['caps'].data_ptr():
            e['caps'].copy_(captions, non_blocking=True)
        if cap_lens.data_ptr() != e['lens'].data_ptr():
            e['lens'].copy_(cap_lens, non_blocking=True)
        e['graph'].replay()
        return e['out']

    @staticmethod
    def _precision():
        from . import get_precision
        return get_precision()

    @staticmethod
    def _graph_key(lr_shape, cap_shape, tmax):
        """Graph signature: shapes, precision and the weights epoch (a graph replays the packed weights it captured;
        training steps or raw-pointer writes since then make it stale, see _lib.bump_weights_epoch)."""
        from . import get_precision
        from ._lib import weights_epoch
        return (tuple(lr_shape), tuple(cap_shape), int(tmax), get_precision(), weights_epoch())

    def static_inputs(self, B: int, tmax: int, Tcap: int = 18, size: int = 32):
        """The static input buffers of the graph for this signature (fill them in place to skip copies)."""
        key = self._graph_key((B, 3, size, size), (B, Tcap), tmax)
        e = getattr(self, '_graphs', {}).get(key)
        return None if e is None else (e['LR'], e['caps'], e['lens'])

    def _capture(self, key, LR, captions, cap_lens, tmax):
        if not hasattr(self, '_graphs'):
            self._graphs = {}
        # the three static inputs are views of ONE device buffer, so a host batch packed the same way needs one H2D copy
        B, size, Tcap = LR.shape[0], LR.shape[2], captions.shape[1]
        blob = torch.empty(self.packed_nbytes(B, Tcap, size), device=self.device, dtype=torch.uint8)
        vLR, vcaps, vlens = self.packed_views(blob, B, Tcap, size)
        vLR.copy_(LR); vcaps.copy_(captions.to(torch.int64)); vlens.copy_(cap_lens.to(device=self.device, dtype=torch.int64))
        ent = {'blob': blob, 'LR': vLR, 'caps': vcaps, 'lens': vlens}
        cur = torch.cuda.current_stream(self.device)
        side = torch.cuda.Stream(self.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(2):                 # warm every weight-prep cache (BN fold, packed operands, LSTM table)
                self.text_and_generate(ent['LR'], ent['caps'], ent['lens'], tmax)
        cur.wait_stream(side)
        torch.cuda.synchronize(self.device)
        from . import _lib
        lib = _lib.load()
        n0 = lib.tgsr_launch_count()
        g = torch.cuda.CUDAGraph()
        # TGSR_GRAPH_PRIO=1 captures the critical chain (LSTM -> stage 1 -> 2 -> 3 -> NetG_high tail) from a HIGH-priority stream
        # and the side chains from default-priority ones (kernel nodes inherit it).  Measured: stage 1 starts 7 us earlier and
        # stage 3 ends 20 us earlier, but the deferred 32x32 trunk then overlaps the large kernels and the step gets SLOWER
        # (0.730 vs 0.723 ms, profiles/r2_experiments.md) -- off by default.
        import os
        prio = int(os.environ.get('TGSR_GRAPH_PRIO', '0'))
        cap = torch.cuda.Stream(self.device, priority=-1) if prio else torch.cuda.Stream(self.device)
        with torch.cuda.graph(g, stream=cap):
            ent['out'] = self.text_and_generate(ent['LR'], ent['caps'], ent['lens'], tmax)
        ent['graph'] = g
        ent['n_kernels'] = int(lib.tgsr_launch_count() - n0)     # libtgsr kernels inside one replay
        self.last_graph_kernels = ent['n_kernels']
        self._graphs[key] = ent
        return ent

    # ------------------------------------------------------------------ packed inputs (one H2D per batch)
    @staticmethod
    def _packed_layout(B: int, Tcap: int, size: int):
        n_lr = B * 3 * size * size * 4
        o_caps = (n_lr + 15) // 16 * 16
        o_lens = (o_caps + B * Tcap * 8 + 15) // 16 * 16
        return n_lr, o_caps, o_lens, o_lens + B * 8

    @classmethod
    def packed_nbytes(cls, B: int, Tcap: int = 18, size: int = 32) -> int:
        return cls._packed_layout(B, Tcap, size)[3]

    @classmethod
    def packed_views(cls, blob: torch.Tensor, B: int, Tcap: int = 18, size: int = 32):
        """(LR fp32 [B,3,size,size], captions int64 [B,Tcap], cap_lens int64 [B]) as views of a uint8 buffer of
        packed_nbytes bytes (host pinned or device).  A host batch built in such a buffer goes to the device in one copy."""
        n_lr, o_caps, o_lens, _ = cls._packed_layout(B, Tcap, size)
        return (blob[:n_lr].view(torch.float32).view(B, 3, size, size),
                blob[o_caps:o_caps + B * Tcap * 8].view(torch.int64).view(B, Tcap),
                blob[o_lens:o_lens + B * 8].view(torch.int64))

    # ------------------------------------------------------------------ roofline probes (bench.py)
    @staticmethod
    def _time_kernel(fn, flush, iters=20, warm=3):
        from .timing import cold_time_ms
        return cold_time_ms(fn, flush, iters=iters, warm=warm)

    @torch.no_grad()
    def dominant_kernel_roofline(self, B, flush, pk):
        """The dominant kernel of the step: ResBlock conv1 of h_net3 (conv3x3 64->128 + BN + GLU at
        128x128; 2*9*64*128*128^2 FLOP per image, util.py:115-117), timed alone."""
        from . import get_precision, is_16bit
        from .blocks import _conv
        from ._lib import ACT_GLU
        blk = self.netGL.h_net3.residual[0].block
        H = W = 128
        x = torch.randn(B, 64, H, W, device=self.device) * 0.5
        flops = 2.0 * 9 * 64 * 128 * H * W * B
        if is_16bit():
            from . import fast, tc, fmt_for
            from .blocks import _folded_bn
            xb = tc.to_blocked(x, fmt=fmt_for('low'))
            yb = tc.Blocked.empty(B, 64, H, W, self.device, fmt=xb.fmt)
            fn = lambda: fast.conv_bn_act(xb, blk[0], blk[1], ACT_GLU, out=yb)
            kernel = 'conv3x3_tc_kernel<64,128,CONV,GLU> (tcgen05, %s operands)' % ('fp16' if xb.fmt else 'bf16')
        else:
            fn, kernel = (lambda: _conv(x, blk[0], blk[1], ACT_GLU)), 'conv2d_f32_kernel<3,GLU> (CUDA-core fp32)'
        ms = self._time_kernel(fn, flush)
        ach = flops / (ms * 1e-3) / 1e12
        return {'bound': 'tensor', 'achieved': ach, 'peak': pk['bf16'], 'unit': 'TFLOP/s', 'frac': ach / pk['bf16'],
                'traffic': _ncu_traffic('conv3x3_tc_kernel<64,128,CONV,GLU> B=32 128x128') if (B == 32 and is_16bit()) else None, 'kernel': kernel, 'ms_per_launch': ms, 'flops_per_launch': flops,
                'peak_source': pk['source'] + ' cuBLAS bf16 burst (kernel timed alone)'}

    @torch.no_grad()
    def kernel_rooflines(self, B, flush, pk, iters: int = 10):
        """Every conv / head kernel variant of the step at its largest instance (and the LSTM), each timed alone, L2-cold:
        achieved TFLOP/s on EXECUTED MMA FLOPs vs the measured bf16 burst peak and algorithmic GB/s vs the measured copy
        bandwidth (whichever bound is tighter names the kernel's roofline).  16-bit modes only."""
        from . import fast, tc, fmt_for, is_16bit
        from ._lib import ACT_GLU, ACT_NONE
        if not is_16bit():
            return None
        dev = self.device
        gl, gh = self.netGL, self.netGH
        st, rb, hb = gl.h_net3, gl.h_net3.residual[0].block, gh.residual48
        fl, fh = fmt_for('low'), fmt_for('high')
        S = 128
        px = B * S * S
        x64 = tc.to_blocked(torch.randn(B, 64, S, S, device=dev) * 0.5, fmt=fl)
        r64 = tc.to_blocked(torch.randn(B, 64, S, S, device=dev) * 0.5, fmt=fl)
        o64 = tc.Blocked.empty(B, 64, S, S, dev, fmt=fl)
        o32u = tc.Blocked.empty(B, 32, 2 * S, 2 * S, dev, fmt=fl)
        x32 = tc.to_blocked(torch.randn(B, 32, S, S, device=dev) * 0.5, fmt=fh)
        r32 = tc.to_blocked(torch.randn(B, 32, S, S, device=dev) * 0.5, fmt=fh)
        o32 = tc.Blocked.empty(B, 32, S, S, dev, fmt=fh)
        o32uh = tc.Blocked.empty(B, 32, 2 * S, 2 * S, dev, fmt=fh)
        add = torch.randn(B, 3, 2 * S, 2 * S, device=dev)
        caps = torch.randint(1, 5450, (B, 18), device=dev); lens = torch.full((B,), 18, device=dev)
        e = 2
        rows = [
            ('conv3x3 64->128 +BN+GLU @128^2 (ResBlock conv 1)', lambda: fast.conv_bn_act(x64, rb[0], rb[1], ACT_GLU, out=o64),
             2 * 9 * 64 * 128 * px, px * (64 + 64) * e),
            ('conv3x3 64->64 +BN+skip @128^2 (ResBlock conv 2)', lambda: fast.conv_bn_act(x64, rb[3], rb[4], ACT_NONE, residual=r64, out=o64),
             2 * 9 * 64 * 64 * px, px * (64 + 64 + 64) * e),
            ('upBlock 64->32 @128^2->256^2 (4-phase)', lambda: fast.conv_bn_act(x64, st.upsample[1], st.upsample[2], ACT_GLU, up2x=True, out=o32u),
             2 * 16 * 64 * 64 * px, px * 64 * e + 4 * px * 32 * e),
            ('conv3x3 32->64 +BN+GLU @128^2 (NetG_high body)', lambda: fast.conv_bn_act(x32, hb[0], hb[1], ACT_GLU, out=o32),
             2 * 9 * 32 * 64 * px, px * (32 + 32) * e),
            ('conv3x3 32->32 +BN+skip @128^2 (NetG_high ResBlock conv 2)', lambda: fast.conv_bn_act(x32, hb[3], hb[4], ACT_NONE, residual=r32, out=o32),
             2 * 9 * 32 * 32 * px, px * (32 + 32 + 32) * e),
            ('upBlock 32->32 @128^2->256^2 (4-phase)', lambda: fast.conv_bn_act(x32, gh.upscale8x[1], gh.upscale8x[2], ACT_GLU, up2x=True, out=o32uh),
             2 * 16 * 32 * 64 * px, px * 32 * e + 4 * px * 32 * e),
            ('image head 3x3 32->3 @256^2', lambda: fast.image_head(o32u, gl.img_net3.img),
             2 * 9 * 32 * 3 * 4 * px, 4 * px * (32 * e + 12)),
            ('image head 5x5 32->3 + tanh + 0.5*SRb @256^2', lambda: fast.image_head(o32uh, gh.conv_output, add, 0.5, 1.0),
             2 * 25 * 32 * 3 * 4 * px, 4 * px * (32 * e + 12 + 12)),
            ('bi-LSTM text encoder, 18 steps', lambda: self.text_encoder._forward_tmax(caps, lens, 18), 15.78e6 * B, None),
        ]
        out = []
        for name, fn, flops, nbytes in rows:
            ms = self._time_kernel(fn, flush, iters=iters)
            tf = flops / (ms * 1e-3) / 1e12
            row = {'kernel': name, 'ms': ms, 'tflops': tf, 'frac_tensor': tf / pk['bf16']}
            if nbytes:
                gbs = nbytes / (ms * 1e-3) / 1e9
                row.update({'gbs': gbs, 'frac_hbm': gbs / pk['hbm']})
            out.append(row)
        return out

    @staticmethod
    @torch.no_grad()
    def attention_sweep(pk, flush, device, quick: bool = True):
        """BASELINE.json configs[2] inside the bench line: GlobalAttentionGeneral(idf 32, cdf 256) at Q in {64^2, 128^2} x
        T in {18, 32, 48}, module API (fp32 NCHW I/O) and the 16-bit blocked path (word projection included), batch
        chosen so that at least 256 MB move; algorithmic bytes / L2-cold device time vs the measured copy
        bandwidth.  scripts/bench_attention.py is the full sweep (other widths, func_attention)."""
        from . import GlobalAttentionGeneral, fast, tc
        from .timing import cold_time_ms
        idf, cdf = 32, 256
        att = GlobalAttentionGeneral(idf, cdf).to(device)
        att.conv_context.weight.data.normal_(0, 0.1)
        rows = []
        for T in (18, 32, 48):
            for S in ((64, 128) if quick else (32, 64, 128, 256)):
                Q = S * S
                for variant, e in (('fp32', 4), ('blocked16', 2)):
                    per_img = Q * (idf * e + idf * e + 4 * T) + cdf * T * 4 + T
                    B = max(1, min(4096, -(-256_000_000 // per_img)))
                    ctx = torch.randn(B, cdf, T, device=device) * 0.3
                    mask = torch.zeros(B, T, dtype=torch.uint8, device=device)
                    if variant == 'fp32':
                        x = torch.randn(B, idf, S, S, device=device) * 0.5
                        att.applyMask(mask)
                        fn = lambda: att(x, ctx)
                    else:
                        cat = tc.Blocked.empty(B, 2 * idf, S, S, device); cat.buf.normal_(0, 0.5)
                        fn = lambda: fast.attention(cat.slice(0, idf), att, ctx, mask, out=cat.slice(idf, idf))
                    ms = cold_time_ms(fn, flush, iters=6, reps=2)
                    nbytes = B * per_img + idf * cdf * 4
                    rows.append({'variant': variant, 'B': B, 'Q': Q, 'T': T, 'idf': idf, 'cdf': cdf, 'us': ms * 1e3,
                                 'gbs': nbytes / ms / 1e6, 'frac_hbm': nbytes / ms / 1e6 / pk['hbm']})
                    del ctx, mask, fn
                    if variant == 'fp32':
                        del x
                    else:
                        del cat
        return rows

    @staticmethod
    def step_roofline_bound(B: int, pk, T: int = 18):
        """Analytic lower bound of one x8 step with ONE kernel per layer (no cross-layer fusion): for every conv / head /
        attention layer max(executed MMA FLOP / measured bf16 peak, algorithmic bytes / measured copy bandwidth), summed
        along the dependent chain (the text encoder and NetG_high's 32x32 trunk run beside G_SR_NET_low's first stages and
        are left out).  `ms_per_step / bound_ms` says how far the whole step is from the sum of its per-kernel rooflines."""
        e, tf, bw = 2, pk['bf16'] * 1e12, pk['hbm'] * 1e9
        t = lambda flops, nbytes: max(flops / tf, nbytes / bw)
        conv = lambda ci, co, px, skip: t(2 * 9 * ci * co * px, px * (ci + (co // 2 if not skip else co) + (co if skip else 0)) * e)
        up = lambda ci, co, px: t(2 * 16 * ci * 2 * co * px, px * ci * e + 4 * px * co * e)
        head = lambda k, px, add: t(2 * k * k * 32 * 3 * px, px * (32 * e + 12 + (12 if add else 0)))
        attn = lambda px: t(0, px * (32 * e + 32 * e + 4 * T))
        total = 0.0
        for S in (32, 64, 128):                                   # G_SR_NET_low stages 1..3 (64 channels, 2 ResBlocks)
            px = B * S * S
            total += attn(px) + 2 * (conv(64, 128, px, False) + conv(64, 64, px, True)) + up(64, 32, px) + head(3, 4 * px, False)
        for S in (64, 128):                                       # NetG_high tail after the trunk (32 channels)
            px = B * S * S
            total += conv(32, 64, px, False) + t(2 * 9 * 32 * 32 * px, px * 64 * e) + up(32, 32, px) + head(5, 4 * px, True)
        total += head(5, B * 64 * 64, True)
        return total * 1e3

    @torch.no_grad()
    def attention_roofline(self, B, flush, pk):
        """Fused word attention at the h_net3 shape (Q=128^2, idf 32, cdf 256, T 18), timed alone.
        Algorithmic bytes per SURVEY.md 8d: B*Q*(idf*e_in + idf*e_out + 4*T) + B*cdf*T*4 + idf*cdf*4 + B*T."""
        from . import get_precision, is_16bit
        att = self.netGL.h_net3.att
        Q, idf, cdf, T = 128 * 128, 32, 256, 18
        ctx = torch.randn(B, cdf, T, device=self.device) * 0.3
        mask = torch.zeros(B, T, dtype=torch.bool, device=self.device)
        att.applyMask(mask)
        if is_16bit():
            from . import fast, tc, fmt_for
            cat = tc.Blocked.empty(B, 64, 128, 128, self.device, fmt=fmt_for('low'))
            cat.buf.normal_(0, 0.5)
            m8 = mask.to(torch.uint8)
            sfrag = fast.attention_project(att, ctx, cat.fmt)        # per-caption operand image, as inside a forward
            fn = lambda: fast.attention(cat.slice(0, 32), att, ctx, m8, out=cat.slice(32, 32), sfrag=sfrag)
            e, kernel = 2, 'attn_blocked_kernel (16-bit blocked activations, fp32 maps; word projection precomputed per caption)'
        else:
            x = torch.randn(B, idf, 128, 128, device=self.device) * 0.5
            fn, e, kernel = (lambda: att(x, ctx)), 4, 'attn_fwd_f32_kernel (fp32 NCHW I/O)'
        nbytes = B * Q * (idf * e + idf * e + 4 * T) + B * cdf * T * 4 + idf * cdf * 4 + B * T
        ms = self._time_kernel(fn, flush)
        ach = nbytes / (ms * 1e-3) / 1e9
        return {'bound': 'hbm', 'achieved': ach, 'peak': pk['hbm'], 'unit': 'GB/s', 'frac': ach / pk['hbm'],
                'traffic': _ncu_traffic('attn_blocked_kernel B=32 128x128 T=18') if (B == 32 and is_16bit()) else None, 'kernel': kernel, 'ms_per_launch': ms, 'bytes_per_launch': nbytes,
                'peak_source': pk['source'] + ' copy bandwidth'}

    def _pin(self, name, t):
        buf = self._pinned.get(name)
        if buf is None or buf.shape != t.shape or buf.dtype != t.dtype:
            buf = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            self._pinned[name] = buf
        buf.copy_(t)
        return buf

    # ------------------------------------------------------------------ pipelined host API
    @torch.no_grad()
    def submit(self, LR_host: torch.Tensor, captions_host: torch.Tensor, cap_lens_host: torch.Tensor,
               out: str = 'uint8', packed: Optional[torch.Tensor] = None) -> int:
        """Enqueue one batch from pinned host buffers and return a ticket for `result`.  Two output slots
        alternate: the device->host copy of batch i runs on a copy stream while batch i+1 computes.
        out='uint8': the saved-image format of trainer_objective.py:153-155, quantised on the device
        ([B,3,H,W] uint8); out='fp32': the raw SR tensor.  `packed`: the pinned uint8 buffer the three host tensors are
        views of (packed_views): the batch then crosses PCIe as one copy instead of three, on a dedicated stream, while the
        previous batch computes.  Host buffers must stay unchanged until `result(ticket)` of this submit returns."""
        tmax = int(cap_lens_host.max())                                    # host-side, no device sync
        B, size, Tcap = LR_host.shape[0], LR_host.shape[2], captions_host.shape[1]
        st = self.static_inputs(B, tmax, Tcap, size)
        if st is None:
            self.forward_graph(LR_host.to(self.device), captions_host.to(self.device), cap_lens_host.to(self.device), tmax)
            st = self.static_inputs(B, tmax, Tcap, size)
        cur = torch.cuda.current_stream(self.device)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        ent = self._graphs[self._graph_key((B, 3, size, size), (B, Tcap), tmax)]
        if packed is not None and packed.numel() == ent['blob'].numel():
            # One H2D for the whole batch, issued on its own stream into one of two staging blobs so that it crosses PCIe
            # while the previous batch still computes; the compute stream then moves it into the graph's inputs with a
            # device-to-device copy (a few microseconds) once that batch has finished.
            if self._h2d_stream is None:
                self._h2d_stream = torch.cuda.Stream(self.device)
            stage = ent.get('stage')
            if stage is None:
                stage = ent['stage'] = [dict(buf=torch.empty_like(ent['blob']), filled=torch.cuda.Event(),
                                             consumed=torch.cuda.Event()) for _ in range(2)]
                for g in stage:
                    g['consumed'].record(cur)
            g = stage[(self._ticket + 1) & 1]
            with torch.cuda.stream(self._h2d_stream):
                self._h2d_stream.wait_event(g['consumed'])               # its previous contents have been moved on
                g['buf'].copy_(packed, non_blocking=True)
                g['filled'].record(self._h2d_stream)
            cur.wait_event(g['filled'])
            ent['blob'].copy_(g['buf'], non_blocking=True)
            g['consumed'].record(cur)
        else:
            st[0].copy_(LR_host, non_blocking=True)
            st[1].copy_(captions_host, non_blocking=True)
            st[2].copy_(cap_lens_host, non_blocking=True)
        sr = self.forward_graph(st[0], st[1], st[2], tmax)[0][-1]
        key = (tuple(sr.shape), out)
        slots = self._slots.get(key)
        if slots is None:
            slots = []
            for _ in range(2):
                s = _Slot()
                dt = torch.uint8 if out == 'uint8' else torch.float32
                s.dev = torch.empty(sr.shape, device=self.device, dtype=dt)
                s.host = torch.empty(sr.shape, dtype=dt, pin_memory=True)
                s.done = torch.cuda.Event(); s.copied = torch.cuda.Event()
                s.flag_host = torch.zeros(1, dtype=torch.int32, pin_memory=True) if self._precision() != 'fp32' else None
                s.copied.record(cur)
                slots.append(s)
            self._slots[key] = slots
        self._ticket += 1
        s = slots[self._ticket & 1]
        cur.wait_event(s.copied)                                           # the slot's previous copy has left the device
        if out == 'uint8':
            from . import ops
            ops.quantize_u8(sr, s.dev)
        else:
            s.dev.copy_(sr, non_blocking=True)
        s.done.record(cur)
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(s.done)
            s.host.copy_(s.dev, non_blocking=True)
            if s.flag_host is not None:
                from . import fast
                s.flag_host.copy_(fast.nonfinite_flag(self.device), non_blocking=True)
            s.copied.record(self._copy_stream)
        self._tickets[self._ticket] = s
        return self._ticket

    def result(self, ticket: int) -> torch.Tensor:
        """Block until the batch of `ticket` is in host memory; the buffer is reused two submits later.  Raises
        FloatingPointError when the 16-bit forward of that batch overflowed (inf / NaN reached an image head)."""
        s = self._tickets.pop(ticket)
        s.copied.synchronize()
        if s.flag_host is not None and int(s.flag_host[0]) != 0:
            from . import fast
            fast.nonfinite_flag(self.device).zero_()
            raise FloatingPointError('tgsr_b200: inf/NaN in the 16-bit forward (an activation exceeded the fp16 / bf16 range); '
                                     'use set_precision("fp32") or "bf16" for this checkpoint')
        return s.host

    @torch.no_grad()
    def forward_host(self, LR_host: torch.Tensor, captions_host: torch.Tensor, cap_lens_host: torch.Tensor,
                     out_host: Optional[torch.Tensor] = None, use_graph: bool = True) -> torch.Tensor:
        """Host buffers in, final SR image (fine_im[-1]) out on the host: H2D of the inputs, the whole
        forward, D2H of the result, all on the current stream (synchronised before returning)."""
        tmax = int(cap_lens_host.max())                                    # host-side, no device sync
        if use_graph:
            st = self.static_inputs(LR_host.shape[0], tmax, captions_host.shape[1], LR_host.shape[2])
            if st is None:
                self.forward_graph(LR_host.to(self.device), captions_host.to(self.device),
                                   cap_lens_host.to(self.device), tmax)
                st = self.static_inputs(LR_host.shape[0], tmax, captions_host.shape[1], LR_host.shape[2])
            st[0].copy_(LR_host, non_blocking=True)                        # H2D straight into the graph's inputs
            st[1].copy_(captions_host, non_blocking=True)
            st[2].copy_(cap_lens_host, non_blocking=True)
            fine_im = self.forward_graph(st[0], st[1], st[2], tmax)[0]
        else:
            LR = LR_host.to(self.device, non_blocking=True)
            caps = captions_host.to(self.device, non_blocking=True)
            lens = cap_lens_host.to(self.device, non_blocking=True)
            fine_im = self.text_and_generate(LR, caps, lens, tmax)[0]
        sr = fine_im[-1]
        if out_host is None:
            out_host = torch.empty(sr.shape, dtype=sr.dtype, pin_memory=True)
        out_host.copy_(sr, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out_host


class _Slot:
    __slots__ = ('dev', 'host', 'done', 'copied', 'flag_host')


def _ncu_traffic(key: str):
    """DRAM bytes per launch of a kernel from the committed `ncu --set full` capture (profiles/ncu_traffic.json)."""
    import json, os
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles', 'ncu_traffic.json')
    try:
        e = json.load(open(p))[key]
        return e['read'] + e['write']
    except (OSError, KeyError, ValueError):
        return None


def randomize_bn_stats(modules, seed: int):
    """SURVEY.md 8d cfg 2: running_mean~N(0,0.1), running_var~U(0.5,1.5), gamma~N(1,0.02), beta=0."""
    import torch.nn as nn
    g = torch.Generator().manual_seed(1234 + seed)
    for mod in modules:
        for m in mod.modules():
            if isinstance(m, (nn.BatchNorm2d, nn.BatchNorm1d)):
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
                m.weight.data.copy_(1 + 0.02 * torch.randn(m.weight.shape, generator=g))
                m.bias.data.zero_()
