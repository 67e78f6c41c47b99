#!/usr/bin/env python
"""bench.py -- SR images/sec of the TGSR eval forward (text encoder -> G_SR_NET_low ->
NetG_highweight) on B200, per BASELINE.json.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision bf16|fp32]

A step is one pass of the hot path over one batch of synthetic input (configs[1]: bird SR eval,
batch 32 per GPU, 32x32 LR -> 256x256, 18-word captions, ntoken 5450, random-init weights with
randomised BN statistics).  Multi-GPU = one process per GPU (torchrun), images sharded by rank,
no collective on the data path (weak scaling); time = max over ranks, measured with CUDA events.

Prints ONE JSON line on rank 0 (see the contract in the task statement):
  value     : images/s, inputs resident in HBM
  e2e       : same metric through SRPipeline.forward_host (pinned host buffers, H2D + D2H inside)
  roofline  : the dominant conv kernel timed alone with CUDA events vs the measured bf16 peak
  cpu_baseline : the reference's own modules (baseline/_ref, unmodified) on this box's host cores
  sustained : >= 3 s of back-to-back replays with their own clock record; roofline_per_kernel: every conv variant timed alone
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
if os.environ.get('NCCL_DEBUG', 'VERSION').upper() in ('VERSION', 'WARN'):
    os.environ['NCCL_DEBUG'] = 'NONE'      # keep stdout to the single JSON line (NCCL prints its version there from VERSION up)

BATCH_PER_GPU = 32
NTOKEN = 5450
WORDS = 18
GFLOP_PER_IMAGE = 21.12          # reference FLOP count, 2*MAC, x8 nets
METRIC = 'SR images/sec (eval fwd, x8 32->256, batch 32/GPU)'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--precision', default=None, choices=['fp32', 'bf16', 'fp16', 'mixed16'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--global-batch', type=int, default=0,
                    help='configs[3] (batch 512 sharded): total images per step, split evenly over the ranks (strong scaling); '
                         '0 = the headline config, 32 images per GPU (weak scaling)')
    ap.add_argument('--no-graph', action='store_true', help='launch every kernel from Python instead of CUDA-graph replay')
    ap.add_argument('--no-kernel-table', action='store_true', help='skip the per-kernel roofline table (roofline_per_kernel)')
    ap.add_argument('--sustained-s', type=float, default=3.0, help='seconds of back-to-back replays for the `sustained` leg (0 = skip)')
    ap.add_argument('--no-train', action='store_true', help='skip the `train` block (configs[4]: 3 timed generator training steps)')
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm=d['hbm_gbs'], bf16=d['bf16_tflops'], bf16_sustained=d.get('bf16_tflops_sustained', d['bf16_tflops']),
                    source='measured')
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source='fallback')


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML, every 5 ms, own thread)."""

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self.t = None
        self.err = None

    def _run(self):
        try:
            import pynvml as N
            N.nvmlInit()
            h = N.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))
            names = {'hw_slowdown': N.nvmlClocksEventReasonHwSlowdown if hasattr(N, 'nvmlClocksEventReasonHwSlowdown') else 0x8,
                     'hw_thermal_slowdown': 0x40, 'sw_thermal_slowdown': 0x20, 'sw_power_cap': 0x4}
            get_reasons = getattr(N, 'nvmlDeviceGetCurrentClocksEventReasons', None) or \
                getattr(N, 'nvmlDeviceGetCurrentClocksThrottleReasons')
            while not self._stop.is_set():
                self.samples.append(float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)))
                r = int(get_reasons(h))
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
                time.sleep(0.005)
        except Exception as e:                      # noqa: BLE001
            self.err = repr(e)

    def start(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def clear(self):
        self.samples.clear(); self.reasons.clear()

    def stop(self):
        self._stop.set()
        if self.t is not None:
            self.t.join(timeout=2)
        sm = sorted(self.samples)
        out = {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
               'samples': len(sm)}
        if self.err:
            out['error'] = self.err
        return out


# ----------------------------------------------------------------------------- reference arm
def workload_config(B, world, precision=None, launch=None):
    cfg = {'workload': 'bird SR eval (cfg/eval_birdSR_attn2.yml dims): text encoder + G_SR_NET_low + '
                       'NetG_highweight, x8 32->256, batch %d per GPU, 18-word captions, ntoken %d, '
                       'random-init weights, randomised BN stats' % (B, NTOKEN),
           'batch_per_gpu': B, 'parallelism': 'image-sharded replicas x%d (no collective)' % world,
           'gflop_per_image_reference': GFLOP_PER_IMAGE}
    if precision is not None:
        cfg.update({'l2': 'flushed between timed iterations (256 MiB write)', 'precision': precision, 'launch': launch})
    return cfg


def cpu_reference_leg(budget_quick: bool = True):
    """`cpu_baseline` of the product arm: the reference's OWN modules (baseline/_ref, unmodified) on this box's host cores,
    SURVEY.md 8d protocol (B = 1 and 8, all threads and one thread, 3 warm-ups, median of >= 10 passes); ~20 s.
    Falls back to the CPU oracle port when baseline/_ref was not installed."""
    cores = os.cpu_count() or 1
    from baseline import ref_arm
    if ref_arm.available():
        ref = ref_arm.load()
        nets = ref_arm.build(ref, NTOKEN)
        rows = ref_arm.cpu_protocol(nets, NTOKEN, WORDS, cores, quick=budget_quick)
        r = rows['all_threads_b8']
        return {'value': r['images_per_s'], 'unit': 'images/s', 'cores': cores, 'kind': 'reference',
                'sample': 'unmodified reference modules (baseline/_ref), fp32, %d threads: median of %d passes over one 8-image '
                          'batch of the batch-%d workload after %d warm-ups' % (cores, r['passes'], BATCH_PER_GPU, r['warmup']),
                'protocol': rows}
    v, sample = cpu_oracle_rate(4, 12.0, cores)
    return {'value': v, 'unit': 'images/s', 'cores': cores, 'kind': 'port', 'sample': sample + ' (baseline/_ref not installed)'}


def cpu_oracle_rate(n_images: int, budget_s: float, threads: int):
    """Fallback only: times the CPU oracle (PyTorch fp32 port of the reference path) on `n_images`-image batches."""
    import torch
    from oracle import tgsr_oracle as O
    import tgsr_b200
    tgsr_b200.set_eval_defaults()
    from tgsr_b200 import nets, text_encoder
    torch.set_num_threads(threads)
    torch.manual_seed(100)
    enc = text_encoder.RNN_ENCODER(NTOKEN, nhidden=256).eval()
    gl = nets.G_SR_NET_low().eval(); gh = nets.NetG_highweight(low='lr').eval()   # parameter holders only
    sdE, sdL, sdH = enc.state_dict(), gl.state_dict(), gh.state_dict()
    LR = O.synth_lr(n_images, 32, seed=101)
    caps, lens = O.synth_captions(n_images, NTOKEN, seed=102, words_num=WORDS)

    def one():
        with torch.no_grad():
            words, sent = O.rnn_encoder(sdE, caps, lens)
            mask = (caps == 0)[:, :words.shape[2]]
            imgs, atts, mu, lv, _ = O.g_sr_net_low(sdL, LR, sent, words, mask)
            return O.netg_high(sdH, LR, imgs, LR, low='lr')
    one()                                               # warm-up
    t0 = time.perf_counter(); n = 0
    while n < 2 or time.perf_counter() - t0 < budget_s:
        one(); n += 1
    dt = time.perf_counter() - t0
    return n * n_images / dt, 'CPU oracle port: %d passes of %d images (x8, 32->256, 18 words) in %.1f s' % (n, n_images, dt)


def run_reference(args):
    """The reference arm: the reference's own modules (baseline/_ref) through the gen_exampleSRHL body on the host cores,
    all threads.  K timed steps after W warm-up steps, exactly as asked; a step = one pass over a batch of the product arm's
    workload (32 images; 8 when K > 60 so the run stays within minutes -- stated in `sample`)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    import torch
    from baseline import ref_arm
    steps, warm = args.steps, args.warmup
    n_img = BATCH_PER_GPU if steps <= 60 else 8
    if not ref_arm.available():
        v, sample = cpu_oracle_rate(4, 12.0, cores)
        print(json.dumps({'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'images/s', 'n_gpus': args.gpus,
                          'steps': steps, 'warmup': warm, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                          'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(BATCH_PER_GPU, 1),
                          'cpu_baseline': {'value': v, 'unit': 'images/s', 'cores': cores, 'kind': 'port', 'sample': sample},
                          'e2e': {'value': v, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))
        return
    ref = ref_arm.load()
    nets = ref_arm.build(ref, NTOKEN)
    ts = ref_arm.time_cpu(nets, n_img, NTOKEN, WORDS, cores, warm, steps)
    total = sum(ts)
    value = steps * n_img / total
    sample = ('unmodified reference modules (baseline/_ref), fp32, %d threads: %d timed steps of %d images after %d warm-up '
              'steps' % (cores, steps, n_img, warm))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'images/s', 'n_gpus': args.gpus, 'steps': steps,
        'warmup': warm, 'ms_per_step': 1e3 * total / steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(BATCH_PER_GPU, 1),
        'cpu_baseline': {'value': value, 'unit': 'images/s', 'cores': cores, 'kind': 'reference', 'sample': sample},
        'e2e': {'value': value, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0}
    line['config']['images_per_step'] = n_img
    # SURVEY.md 8d protocol rows (B = 1 / 8, all threads / one thread, median of >= 10), bounded
    line['cpu_protocol'] = ref_arm.cpu_protocol(nets, NTOKEN, WORDS, cores, quick=False)
    # Informational (SURVEY.md 8d "library comparison"): the same unmodified modules under torch eager on the B200
    if torch.cuda.is_available():
        try:
            line['torch_eager_b200'] = ref_arm.torch_eager_gpu(ref, NTOKEN, WORDS, BATCH_PER_GPU)
        except Exception as e:                       # noqa: BLE001  (never let the extra leg break the reference arm)
            line['torch_eager_b200'] = {'error': repr(e)[:200]}
        try:                                         # the same for configs[4]: the generator training step under eager + autograd
            line['torch_eager_b200_train'] = ref_arm.torch_eager_train_gpu(ref, NTOKEN, WORDS, BATCH_PER_GPU)
        except Exception as e:                       # noqa: BLE001
            line['torch_eager_b200_train'] = {'error': repr(e)[:200]}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import tgsr_b200
    from tgsr_b200 import _lib
    from tgsr_b200.pipeline import SRPipeline
    from tgsr_b200 import synth as O              # synthetic-input generators

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device -- the product path has no CPU fallback')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    tgsr_b200.set_eval_defaults()
    precision = args.precision or tgsr_b200.DEFAULT_BENCH_PRECISION
    tgsr_b200.set_precision(precision)

    pipe = SRPipeline(NTOKEN, device=dev, variant='x8', low='lr', seed=100, randomize_bn=True)
    B = BATCH_PER_GPU
    if args.global_batch:
        if args.global_batch % world:
            raise SystemExit('--global-batch must be a multiple of the number of ranks')
        B = args.global_batch // world
    # host batch in ONE pinned buffer (LR | captions | lengths): a step's inputs cross PCIe as a single copy
    packed_h = torch.empty(SRPipeline.packed_nbytes(B, WORDS, 32), dtype=torch.uint8).pin_memory()
    LR_h, caps_h, lens_h = SRPipeline.packed_views(packed_h, B, WORDS, 32)
    LR_h.copy_(O.synth_lr(B, 32, seed=101 + rank))
    c_, l_ = O.synth_captions(B, NTOKEN, seed=102 + rank, words_num=WORDS)
    caps_h.copy_(c_); lens_h.copy_(l_)
    LR, caps, lens = LR_h.to(dev), caps_h.to(dev), lens_h.to(dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev, dtype=torch.float32)     # 256 MiB > 126 MB L2

    use_graph = not args.no_graph
    if use_graph:                      # capture once; the resident inputs become the graph's static inputs
        pipe.forward_graph(LR, caps, lens, WORDS)
        LR, caps, lens = pipe.static_inputs(B, WORDS, WORDS, 32)

    def step_device():
        if use_graph:
            return pipe.forward_graph(LR, caps, lens, WORDS)[0]
        return pipe.text_and_generate(LR, caps, lens, WORDS)[0]

    out_host = torch.empty((B, 3, 256, 256), dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing
    sampler = ClockSampler(int(os.environ.get('CUDA_VISIBLE_DEVICES', '0,1,2,3,4,5,6,7').split(',')[local])
                           if os.environ.get('CUDA_VISIBLE_DEVICES') else local)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler.clear()                     # keep only samples taken under the timed region's load
    lib = _lib.load()
    launches0 = lib.tgsr_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for s, e in ev:
        flush.zero_()                         # L2 flush between timed iterations (outside the event pair)
        s.record(); step_device(); e.record()
    barrier()
    clocks = sampler.stop()
    launches = lib.tgsr_launch_count() - launches0
    if use_graph:                       # replays do not pass through the library's counter
        launches = pipe.last_graph_kernels * args.steps
    t_ms = sum(s.elapsed_time(e) for s, e in ev)
    t = torch.tensor([t_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_ms = float(t.item())
    value = world * B * args.steps / (t_ms * 1e-3)

    # ---- end-to-end timing through the public API with host buffers: SRPipeline.submit / result.
    # Every step copies its inputs from pinned host memory (H2D), runs the whole forward, quantises the SR image to
    # the saved-image format (uint8, trainer_objective.py:153-155) and copies it to pinned host memory (D2H); the H2D
    # of step i+1 (own stream, two staging buffers) and the D2H of step i (two output slots, copy stream) overlap the
    # compute of the neighbouring steps.  Every result is waited for on the
    # host inside the timed region.  No explicit L2 flush here: one step streams ~1.3 GB through the 126 MB L2.
    def e2e_loop(n, out):
        prev = None
        for _ in range(n):
            tk = pipe.submit(LR_h, caps_h, lens_h, out=out, packed=packed_h)
            if prev is not None:
                pipe.result(prev)
            prev = tk
        return pipe.result(prev)

    def e2e_rate(out):
        e2e_loop(3, out)
        barrier()
        s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s2.record()
        res = e2e_loop(args.steps, out)
        e2.record()
        barrier()
        t2 = torch.tensor([s2.elapsed_time(e2)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        return world * B * args.steps / (float(t2.item()) * 1e-3), res

    if use_graph:
        e2e_value, res = e2e_rate('uint8')
        e2e_fp32, res32 = e2e_rate('fp32')
        d2h, d2h_fp32 = res.numel() * res.element_size(), res32.numel() * res32.element_size()
    else:                                   # per-kernel launches: the blocking call, fp32 result
        for _ in range(3):
            pipe.forward_host(LR_h, caps_h, lens_h, out_host, use_graph)
        barrier()
        s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s2.record()
        for _ in range(args.steps):
            pipe.forward_host(LR_h, caps_h, lens_h, out_host, use_graph)
        e2.record()
        barrier()
        t2 = torch.tensor([s2.elapsed_time(e2)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        e2e_value = e2e_fp32 = world * B * args.steps / (float(t2.item()) * 1e-3)
        d2h = d2h_fp32 = out_host.numel() * out_host.element_size()
    h2d = packed_h.numel() if use_graph else LR_h.numel() * 4 + caps_h.numel() * 8 + lens_h.numel() * 8

    # ---- roofline of the dominant kernel (timed alone, CUDA events, L2 flushed between launches)
    pk = peaks()
    roof = pipe.dominant_kernel_roofline(B, flush, pk) if rank == 0 else None
    attn_roof = pipe.attention_roofline(B, flush, pk) if rank == 0 else None
    # the same kernel with the grid filled 3x deeper (96 images: 4.5 waves of 148 CTAs x 4 resident): separates the
    # kernel's own efficiency from the one-and-a-half-wave tail of the batch-32 launch
    attn_roof_b96 = pipe.attention_roofline(96, flush, pk) if (rank == 0 and not args.no_kernel_table) else None

    per_kernel = pipe.kernel_rooflines(B, flush, pk) if (rank == 0 and not args.no_kernel_table) else None
    attn_sweep = SRPipeline.attention_sweep(pk, flush, dev) if (rank == 0 and not args.no_kernel_table) else None
    bound_ms = SRPipeline.step_roofline_bound(B, pk)
    step_roof = {'bound_ms': bound_ms, 'frac': bound_ms / (t_ms / args.steps),
                 'what': 'sum over the layers of the dependent chain of max(executed MMA FLOP / measured bf16 peak, algorithmic bytes / '
                         'measured copy bandwidth), one kernel per layer (no cross-layer fusion)'}

    # ---- sustained leg: >= 3 s of back-to-back replays (no L2 flush, no gaps), own clock record: what a long-running
    # job gets once the power cap (1000 W) has pulled the SM clock down; the headline above is the burst figure.
    sustained = None
    if use_graph and args.sustained_s > 0:
        n_rep = max(50, int(args.sustained_s / (t_ms / args.steps * 1e-3)))
        s_samp = ClockSampler(sampler.index)
        barrier()
        s_samp.start()
        s3, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s3.record()
        for _ in range(n_rep):
            step_device()
        e3.record()
        barrier()
        t3 = torch.tensor([s3.elapsed_time(e3)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t3, op=dist.ReduceOp.MAX)
        sustained = {'value': world * B * n_rep / (float(t3.item()) * 1e-3), 'unit': 'images/s', 'steps': n_rep,
                     'seconds': float(t3.item()) * 1e-3, 'ms_per_step': float(t3.item()) / n_rep, 'clocks': s_samp.stop(),
                     'l2': 'not flushed (a step streams ~1.3 GB through the 126 MB L2)'}

    # ---- configs[4] in the same run, at every N: 3 timed steps of the bf16 generator training step (per-GPU batch 32,
    # NCCL all-reduce of the flat gradient bucket between the two captured graphs when N > 1)
    train_block = None
    if not args.no_train:
        import bench_train
        del flush
        torch.cuda.empty_cache()
        ta = argparse.Namespace(gpus=world, steps=3, warmup=3, batch=BATCH_PER_GPU, no_graph=False, precision='bf16')
        try:
            r = bench_train.run(ta, e2e=False, kernel_roofline=False)
            if r is not None:
                train_block = {k: r[k] for k in ('metric', 'value', 'unit', 'ms_per_step', 'allreduce_ms_last_step', 'steps', 'warmup',
                                                 'dtype', 'gpu_launches', 'loss')}
                train_block['config'] = r['config']
        except Exception as e:                              # noqa: BLE001  (the headline line must survive a failure here)
            train_block = {'error': repr(e)[:300]}
        tgsr_b200.set_precision(precision)

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base = cpu_reference_leg()

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': t_ms / args.steps, 'higher_is_better': True,
            'scaling': 'strong' if args.global_batch else 'weak', 'vs_baseline': None, 'dtype': {'fp32': 'f32', 'bf16': 'bf16', 'fp16': 'f16', 'mixed16': 'bf16+f16 (16-bit tcgen05 operands, f32 accumulate)'}[precision],
            'data': 'synthetic',
            'config': workload_config(B, world, precision, 'CUDA-graph replay' if use_graph else 'per-kernel launches from Python'),
            'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': 'images/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'result': 'final SR image as uint8 [B,3,256,256] (device-side quantisation of trainer_objective.py:153-155); '
                              'the packed H2D copy of step i+1 (own stream, two staging buffers) and the D2H of step i overlap the compute of the '
                              'neighbouring steps; every result waited for on the host',
                    'fp32_result': {'value': e2e_fp32, 'unit': 'images/s', 'd2h_bytes_per_step': d2h_fp32}},
            'gpu_launches': int(launches),
            'roofline': roof, 'roofline_attention': attn_roof, 'roofline_attention_b96': attn_roof_b96, 'attention_sweep': attn_sweep, 'roofline_per_kernel': per_kernel, 'roofline_step': step_roof, 'sustained': sustained, 'train': train_block,
            'model_tflops': value * GFLOP_PER_IMAGE / 1e3, 'peaks': pk,
        }
        if cpu_base is not None:
            line['cpu_baseline'] = cpu_base
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
