#!/usr/bin/env python
"""bench_train.py -- generator training step of BASELINE.json configs[4] (SURVEY.md 8d cfg 5).

    python bench_train.py [--steps K] [--warmup W] [--batch B]          (1 GPU)
    torchrun --nproc-per-node N ... bench_train.py --gpus N ...         (N GPUs, NCCL gradient all-reduce)

One step = G_SR_NET_low + NetG_highweight forward in train() mode (batch-statistics BatchNorm) on a synthetic batch,
DAMSM word/sentence ranking loss on the final 256x256 image (frozen random-init Inception-v3 image encoder and frozen text
encoder) + MSE against the real pyramid, backward, flat-bucket gradient all-reduce, fused Adam(2e-4, (0.5, 0.999)).
fp32 (the training kernels are CUDA-core fp32 in round 1).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
if os.environ.get('NCCL_DEBUG', 'VERSION').upper() in ('VERSION', 'WARN'):
    os.environ['NCCL_DEBUG'] = 'NONE'      # keep stdout to the single JSON line (NCCL prints its version there from VERSION up)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--batch', type=int, default=32, help='images per GPU')
    ap.add_argument('--no-graph', action='store_true', help='launch every kernel from Python instead of replaying the captured step')
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'tf32', 'fp32'],
                    help="conv precision of the step: bf16 = tcgen05 convs on blocked bf16 activations (BASELINE configs[4]), tf32 / fp32 = round-1 paths")
    return ap.parse_args()


def run(a, e2e: bool = True, kernel_roofline: bool = True):
    """One measurement of the training step with the settings in `a` (steps, warmup, batch, no_graph, precision); returns
    the result dict on rank 0 (None elsewhere).  Uses the default process group when one is initialised."""
    import torch
    import torch.distributed as dist
    import tgsr_b200
    from tgsr_b200 import _lib, nets
    from tgsr_b200.text_encoder import RNN_ENCODER
    from tgsr_b200.train_step import FrozenImageEncoder, GeneratorTrainStep
    from tgsr_b200 import synth as O                       # synthetic-input generators

    world = int(os.environ.get('WORLD_SIZE', '1')); rank = int(os.environ.get('RANK', '0')); local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench_train.py: no CUDA device -- the product path has no CPU fallback')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    own_group = world > 1 and not dist.is_initialized()
    if own_group:
        dist.init_process_group('nccl', device_id=dev)
    prev_precision = tgsr_b200.get_precision()
    tgsr_b200.set_eval_defaults(); tgsr_b200.set_precision('fp32')
    import json as _json
    _pk = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    bf16_peak = _json.load(open(_pk))['bf16_tflops'] if os.path.isfile(_pk) else 1590.0
    from tgsr_b200.config import cfg
    cfg.TRAIN.SMOOTH.GAMMA1, cfg.TRAIN.SMOOTH.GAMMA2, cfg.TRAIN.SMOOTH.GAMMA3 = 4.0, 5.0, 10.0
    torch.manual_seed(100)                                  # identical replicas
    gl = nets.G_SR_NET_low().to(dev); gh = nets.NetG_highweight(low='lr').to(dev)
    enc = RNN_ENCODER(5450, nhidden=256).to(dev).eval()
    img = FrozenImageEncoder().to(dev)
    ts = GeneratorTrainStep(gl, gh, enc, img, conv_precision=a.precision)
    B = a.batch
    LR = O.synth_lr(B, 32, seed=201 + rank).to(dev)
    real = [torch.nn.functional.interpolate(LR, scale_factor=2 ** (k + 1), mode='bicubic').clamp(-1, 1) for k in range(3)]
    caps, lens = O.synth_captions(B, 5450, seed=202 + rank, words_num=18)
    caps, lens = caps.to(dev), lens.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    use_graph = not a.no_graph      # N > 1: two graphs per step with the (uncaptured) NCCL all-reduce between them
    if use_graph:
        ts.capture(LR, real, caps, lens, 18)

    def one(timed=False):
        if use_graph:
            return ts.step_graph(ts._static[0], ts._static[1], ts._static[2], ts._static[3], time_allreduce=timed)
        return ts.step(LR, real, caps, lens, 18, time_allreduce=timed)

    for _ in range(max(a.warmup, 3)):
        out = one()
    barrier()
    lib = _lib.load()
    n0 = lib.tgsr_launch_count()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ar_ms = 0.0
    barrier()
    s.record()
    for _ in range(a.steps):
        out = one(True)
    e.record()
    barrier()
    launches = ts.graph_kernels * a.steps if use_graph else int(lib.tgsr_launch_count() - n0)   # replays bypass the counter
    ar_ms = ts.ev[0].elapsed_time(ts.ev[1]) if ts.ev is not None and world > 1 else float('nan')
    t = torch.tensor([s.elapsed_time(e), ar_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0]) / a.steps
    result = None
    e2e_block = None
    roof = None
    if e2e:
        e2e_block = _e2e_leg(a, ts, use_graph, B, rank, world, dev, caps, lens, barrier, O, dist, torch)
    if rank == 0 and kernel_roofline:
        roof = _conv_roofline(a, B, dev, bf16_peak, torch)
    if rank == 0:
        result = {
            'metric': 'generator training images/sec (G fwd train-BN + DAMSM + MSE + bwd + all-reduce + Adam)',
            'value': world * B / (ms * 1e-3), 'unit': 'images/s', 'n_gpus': world, 'steps': a.steps, 'warmup': max(a.warmup, 3),
            'ms_per_step': ms, 'allreduce_ms_last_step': (None if t[1] != t[1] else float(t[1])), 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': {'bf16': 'bf16 (tcgen05 forward / data-gradient / weight-gradient convs on blocked bf16 activations, fp32 accumulate; fp32 master weights, fp32 BatchNorm statistics)', 'tf32': 'tf32 tensor-core convs (fp32 accumulate), fp32 elsewhere', 'fp32': 'f32'}[a.precision], 'data': 'synthetic',
            'config': {'workload': 'configs[4]: x8 generator training step, batch %d per GPU, 32->256, 18-word captions, '
                                   'gamma (4,5,10), Adam(2e-4,(0.5,0.999)), flat-bucket NCCL all-reduce of %d gradients'
                                   % (B, ts.bucket.numel), 'batch_per_gpu': B, 'launch': ('CUDA-graph replay of the whole step' if world == 1 else 'two CUDA graphs per step (fwd+bwd, Adam) around the eager NCCL all-reduce') if use_graph else 'per-kernel launches from Python',
                       'image_encoder': 'frozen random-init Inception-v3 on torchvision/cuDNN (library, out of scope)' + (', bf16 autocast, channels-last' if a.precision == 'bf16' else '')},
            'gpu_launches': launches, 'roofline': roof,
            'loss': {k: float(v) for k, v in out.items()}}
        if e2e_block is not None:
            result['e2e'] = e2e_block
    tgsr_b200.set_precision(prev_precision)
    if own_group:
        dist.destroy_process_group()
    return result


def _e2e_leg(a, ts, use_graph, B, rank, world, dev, caps, lens, barrier, O, dist, torch):
    # ---- end to end from decoded images: per step the H2D copy of a pinned uint8 [B,256,256,3] batch + captions, the input
    # pyramid on the device (the reference's per-sample PIL work, datasets.py:112-148), the step, and the D2H of the loss.
    hr = torch.nn.functional.interpolate(O.synth_lr(B, 32, seed=201 + rank), size=(256, 256), mode='bicubic').clamp(-1, 1)
    hr_h = ((hr + 1.0) * 127.5).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous().pin_memory()
    caps_h, lens_h = caps.cpu().pin_memory(), lens.cpu().pin_memory()
    loss_h = torch.empty(1, dtype=torch.float32).pin_memory()
    hr_d = torch.empty_like(hr_h, device=dev)
    caps_d, lens_d = torch.empty_like(caps), torch.empty_like(lens)
    if use_graph:
        hr_d.copy_(hr_h); caps_d.copy_(caps_h); lens_d.copy_(lens_h)
        ts.capture_from_images(hr_d, caps_d, lens_d, 18)

    def one_e2e():
        hr_d.copy_(hr_h, non_blocking=True); caps_d.copy_(caps_h, non_blocking=True); lens_d.copy_(lens_h, non_blocking=True)
        o = ts.step_graph_from_images(hr_d, caps_d, lens_d) if use_graph else ts.step_from_images(hr_d, caps_d, lens_d, 18)
        loss_h.copy_(o['total'].reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()                      # the host reads the loss every step
        return float(loss_h[0])
    for _ in range(3):
        one_e2e()
    barrier()
    s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s2.record()
    for _ in range(a.steps):
        last_loss = one_e2e()
    e2.record()
    barrier()
    t2 = torch.tensor([s2.elapsed_time(e2)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_ms = float(t2[0]) / a.steps
    return {'value': world * B / (e2e_ms * 1e-3), 'unit': 'images/s', 'ms_per_step': e2e_ms,
            'h2d_bytes_per_step': hr_h.numel() + caps_h.numel() * 8 + lens_h.numel() * 8, 'd2h_bytes_per_step': 4,
            'input': 'pinned uint8 [B,256,256,3] decoded images + int64 captions; input pyramid (PIL-exact resize, '
                     'normalise) on the device inside the step', 'loss': last_loss}


def _conv_roofline(a, B, dev, bf16_peak, torch):
    roof = None
    if True:
        # dominant library kernel of the step: the TF32 implicit-GEMM conv (64 -> 128 at 128x128), timed alone, L2-cold
        from tgsr_b200 import train
        from tgsr_b200.timing import cold_time_ms, make_flush
        xq = torch.randn(B, 64, 128, 128, device=dev) * 0.5
        wq = torch.randn(128, 64, 3, 3, device=dev) * 0.05
        fl = 2.0 * 9 * 64 * 128 * 128 * 128 * B
        if a.precision == 'bf16':
            from tgsr_b200 import train16
            xb = train16.to_blocked(xq)
            cms = cold_time_ms(lambda: train16._conv_tc(xb, 0, 64, wq), make_flush(dev), iters=5)
            ach = fl / (cms * 1e-3) / 1e12
            roof = {'bound': 'tensor', 'achieved': ach, 'peak': bf16_peak, 'unit': 'TFLOP/s', 'frac': ach / bf16_peak, 'traffic': None,
                    'kernel': 'weight pack + conv3x3_tc_kernel<64,128,CONV,none> (tcgen05, bf16 operands, fp32 accumulate): the forward '
                              'conv of ResBlock conv 1 at 128x128; the data gradient runs the same kernel',
                    'ms_per_launch': cms, 'flops_per_launch': fl, 'peak_source': 'measured cuBLAS bf16 burst (MEASURED_PEAKS.json)'}
        cms = cold_time_ms(lambda: train._conv_raw(xq, wq, False), make_flush(dev), iters=5)
        ach = fl / (cms * 1e-3) / 1e12
        roof_tf32 = {'bound': 'tensor', 'achieved': ach, 'peak': 275.0, 'unit': 'TFLOP/s', 'frac': ach / 275.0, 'traffic': None,
                'kernel': 'conv_prep_tf32_kernel + conv_tf32_kernel<3,64> (mma.sync m16n8k8 TF32, fp32 accumulate)', 'ms_per_launch': cms,
                'flops_per_launch': fl, 'peak_source': 'measured HMMA TF32 issue rate on B200 (scripts/micro/hmma_rate.cu, profiles/r1_hmma_rate.md); '
                                                       'these kernels still compute the weight gradients and the tf32 / fp32 modes'}
        if a.precision != 'bf16':
            roof = roof_tf32
        del xq, wq
    return roof


def main():
    a = parse()
    res = run(a)
    if res is not None:
        print(json.dumps(res))


if __name__ == '__main__':
    main()
