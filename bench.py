#!/usr/bin/env python
"""bench.py -- PIDNet hot-path benchmark (contract: see DESIGN.md "Measurement").

  python bench.py --gpus N --steps K --warmup W          # our engine; under torchrun for N > 1
  python bench.py --impl reference --steps K --warmup W  # the reference's CPU forward (oracle port)

A "step" is one eval forward of a batch of `--batch` synthetic 3x1024x2048 images per GPU through
the PIDNet-S engine (BASELINE.json configs[1]).  `value` = images/s over all ranks with the inputs
resident in HBM; `e2e` = the same metric through the public API (`PIDNet.forward`) with HOST pinned
inputs and the logits read back to the host every step (H2D/D2H inside the timed region, copies
double-buffered against compute).  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# keep stdout to the single JSON line: NCCL prints its version banner to stdout at NCCL_DEBUG=VERSION
# (and at every higher level), so route NCCL's log to stderr and drop a bare VERSION request
if os.environ.get('NCCL_DEBUG', '').upper() == 'VERSION':
    del os.environ['NCCL_DEBUG']
os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')

import torch  # noqa: E402

METRIC = 'pidnet_s_1024x2048_images_per_sec'
UNIT = 'img/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--model', default='pidnet_s')
    ap.add_argument('--classes', type=int, default=19)
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--height', type=int, default=1024)
    ap.add_argument('--width', type=int, default=2048)
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--skip-cpu-baseline', action='store_true')
    ap.add_argument('--skip-extras', action='store_true', help='skip bs1 latency / e2e / per-kernel profile')
    ap.add_argument('--profile-out', default=None, help='write the per-launch table (json) here')
    return ap.parse_args()


def workload_config(a, world):
    return {
        'workload': f'{a.model} eval forward, synthetic {a.batch}x3x{a.height}x{a.width} fp32 NCHW per GPU -> '
                    f'fp32 logits [{a.batch},{a.classes},{a.height // 8},{a.width // 8}] (BASELINE.json configs[1])',
        'batch_per_gpu': a.batch, 'global_batch': a.batch * world, 'height': a.height, 'width': a.width,
        'classes': a.classes, 'weights': 'random init, seed 0', 'parallelism': f'batch-sharded x{world}, no collective',
        'l2': f'per-step input ({a.batch * 3 * a.height * a.width * 4 / 1e6:.0f} MB) and activations exceed the 126 MB L2; '
              'no explicit flush needed for the throughput loop; the bs1 latency loop flushes L2 between iterations',
    }


# ------------------------------------------------------------------------------------------- CPU arm
def cpu_reference_forward(a, steps, warmup):
    """The reference's CPU forward for this path, restated in oracle/ (torch CPU ops, all host threads).
    One step = ONE image of the workload (bounded sample)."""
    from oracle import pidnet_oracle as O
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    cfg = O.config_for(a.model, a.classes, False)
    sd = O.make_state_dict(cfg, seed=0, randomize_bn=False)
    x = torch.randn(1, 3, a.height, a.width, generator=torch.Generator().manual_seed(0))
    ts = []
    with torch.no_grad():
        for _ in range(warmup):
            O.pidnet_forward(sd, x)
        for _ in range(steps):
            t = time.perf_counter()
            O.pidnet_forward(sd, x)
            ts.append(time.perf_counter() - t)
    cpu_model = ''
    try:
        with open('/proc/cpuinfo') as f:
            for line in f:
                if line.startswith('model name'):
                    cpu_model = line.split(':', 1)[1].strip()
                    break
    except OSError:
        pass
    return dict(value=1.0 / statistics.median(ts), unit=UNIT, cores=cores, kind='port',
                sample=f'{steps} forwards of 1x3x{a.height}x{a.width} fp32 (1 image of the batch), median; '
                       f'oracle/pidnet_oracle.py on torch CPU ({torch.get_num_threads()} threads)',
                cpu=cpu_model, median_s=statistics.median(ts), min_s=min(ts))


def run_reference(a):
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if rank != 0:
        return
    steps = max(1, min(a.steps, 40))
    warmup = max(1, min(a.warmup, 3))
    cb = cpu_reference_forward(a, steps, warmup)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': cb['value'], 'unit': UNIT, 'n_gpus': a.gpus, 'steps': steps,
        'warmup': warmup, 'ms_per_step': cb['median_s'] * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(a, world),
        'cpu_baseline': cb,
        'e2e': {'value': cb['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ('timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')
    NAMES = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                       '-lms', '20', '-i', str(index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def wait_ready(self, timeout=3.0):
        """Block until nvidia-smi has written its first sample (so the loop is running before the timed region)."""
        t_end = time.time() + timeout
        while self.p is not None and time.time() < t_end:
            try:
                if os.path.getsize(self.f.name) > 0:
                    return True
            except OSError:
                pass
            time.sleep(0.01)
        return False

    def stop(self, t0=None, t1=None):
        """Samples whose nvidia-smi timestamp lies inside the timed region [t0, t1] (host epoch seconds); the sampler is
        started before the warm-up so that it is already running when the region begins.  If the region was too short to
        catch one, the samples taken under load just before it (warm-up) are reported and `window` says so."""
        import datetime
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        rows = []
        for line in self.f.read().splitlines():
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], '%Y/%m/%d %H:%M:%S.%f').timestamp()
                rows.append((ts, float(parts[1]), float(parts[2]), float(parts[3]),
                             [nm for nm, v in zip(self.NAMES, parts[4:8]) if v.lower().startswith('active')]))
            except ValueError:
                continue
        self.f.close()
        os.unlink(self.f.name)
        window = 'all'
        sel = rows
        if t0 is not None and t1 is not None:
            inside = [r for r in rows if t0 <= r[0] <= t1]
            if inside:
                sel, window = inside, 'timed region'
            else:   # the last samples before the region ended (GPU already under the warm-up load)
                sel, window = [r for r in rows if r[0] <= t1][-3:], 'warm-up + timed region (region shorter than the sampling period)'
        if sel:
            out.update(sm_mhz=statistics.median(r[1] for r in sel), sm_max_mhz=max(r[2] for r in sel),
                       power_w_max=max(r[3] for r in sel), reasons=sorted({n for r in sel for n in r[4]}),
                       samples=len(sel), window=window)
        return out


# ------------------------------------------------------------------------------------------- GPU arm
def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d['hbm_gbs'], tc_burst=d['bf16_tflops'], tc_sustained=d['bf16_tflops_sustained'],
                    source='MEASURED_PEAKS.json (measured)')
    return dict(hbm=6650.0, tc_burst=1590.0, tc_sustained=1400.0, source='B200_PROFILING.md fallback')


def measured_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed `ncu --set full` capture of
    this same command (profiles/r01/traffic.json, written by tools/ncu_traffic.py); None if that kernel was not captured."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'profiles', 'r01', 'traffic.json')
    try:
        with open(path) as f:
            t = json.load(f)
        e = t.get(kernel)
        return None if e is None else e['dram_bytes_per_launch']
    except (OSError, ValueError, KeyError):
        return None


def make_model(a, dev):
    from pidnet_b200 import get_pred_model
    torch.manual_seed(0)
    model = get_pred_model(a.model, a.classes).to(dev).eval()
    return model


def run_ours(a):
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the engine has no CPU fallback); use --impl reference for the CPU arm')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    cpu_base = None
    if rank == 0 and a.gpus == 1 and not a.skip_cpu_baseline:
        cpu_base = cpu_reference_forward(a, steps=5, warmup=2)

    use_graph = not a.no_graph
    B, H, W = a.batch, a.height, a.width
    model = make_model(a, dev)
    g = torch.Generator(device='cpu').manual_seed(1234 + rank)
    x_host = torch.randn(B, 3, H, W, generator=g).pin_memory()
    x = x_host.to(dev, non_blocking=True)
    out = torch.empty(B, a.classes, H // 8, W // 8, device=dev)
    with torch.no_grad():
        sampler = ClockSampler(local) if rank == 0 else None   # started early: nvidia-smi takes ~0.1 s to deliver its first line
        if sampler:
            sampler.wait_ready()
        for _ in range(max(a.warmup, 3)):
            model.forward_into(x, out, use_graph=use_graph)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_host0 = time.time()
        e0.record()
        for _ in range(a.steps):
            model.forward_into(x, out, use_graph=use_graph)
        e1.record()
        barrier()
        t_host1 = time.time()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop(t_host0, t_host1) if sampler else None
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    launches = model.num_launches()
    value = world * B * a.steps / (ms / 1e3)
    flops_per_img = model.conv_flops() / B

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': a.steps, 'warmup': max(a.warmup, 3),
        'ms_per_step': ms / a.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16',
        'data': 'synthetic', 'config': workload_config(a, world), 'clocks': clocks,
        'gpu_launches': launches * a.steps, 'launches_per_step': launches, 'cuda_graph': use_graph,
        'conv_gflop_per_image': flops_per_img / 1e9,
    }
    peaks = measured_peaks()

    if not a.skip_extras:
        # ---- per-launch profile of the same step (ops serialised, CUDA events around each launch)
        with torch.no_grad():
            model.profile(x, out)
            rows = model.profile(x, out)
        groups = {}
        for r in rows:
            gk = groups.setdefault(r['kernel'], dict(kernel=r['kernel'], launches=0, ms=0.0, flops=0.0, bytes=0.0))
            gk['launches'] += 1; gk['ms'] += r['ms']; gk['flops'] += r['flops']; gk['bytes'] += r['bytes']
        tot_ms = sum(gk['ms'] for gk in groups.values())
        top = sorted(groups.values(), key=lambda z: -z['ms'])
        dom = top[0]
        tf = dom['flops'] / (dom['ms'] * 1e-3) / 1e12
        gbs = dom['bytes'] / (dom['ms'] * 1e-3) / 1e9
        tensor_bound = dom['kernel'].startswith('conv')
        line['roofline'] = {
            'kernel': dom['kernel'], 'bound': 'tensor' if tensor_bound else 'hbm',
            'achieved': tf if tensor_bound else gbs, 'peak': peaks['tc_sustained'] if tensor_bound else peaks['hbm'],
            'unit': 'TFLOP/s' if tensor_bound else 'GB/s',
            'frac': (tf / peaks['tc_sustained']) if tensor_bound else (gbs / peaks['hbm']),
            'traffic': measured_traffic(dom['kernel']), 'peak_source': peaks['source'] + (' bf16 sustained' if tensor_bound else ' hbm copy'),
            'launches_per_step': dom['launches'], 'avg_launch_ms': dom['ms'] / dom['launches'],
            'algorithmic_gflop_per_launch': dom['flops'] / dom['launches'] / 1e9,
            'algorithmic_mb_per_launch': dom['bytes'] / dom['launches'] / 1e6,
            'share_of_step': dom['ms'] / tot_ms, 'hbm_gbs_same_kernel': gbs, 'hbm_frac_same_kernel': gbs / peaks['hbm'],
        }
        line['kernels'] = [dict(kernel=k['kernel'], launches=k['launches'], ms=round(k['ms'], 4),
                                share=round(k['ms'] / tot_ms, 4),
                                tflops=round(k['flops'] / (k['ms'] * 1e-3) / 1e12, 2),
                                gbs=round(k['bytes'] / (k['ms'] * 1e-3) / 1e9, 1)) for k in top[:8]]
        line['whole_net'] = {'conv_tflops': flops_per_img * B / (ms / a.steps * 1e-3) / 1e12,
                             'frac_of_tc_sustained': flops_per_img * B / (ms / a.steps * 1e-3) / 1e12 / peaks['tc_sustained'],
                             'serialised_sum_ms': tot_ms}
        if a.profile_out and rank == 0:
            with open(a.profile_out, 'w') as f:
                json.dump(rows, f, indent=1)

        # ---- e2e through the public API: pinned host input -> H2D -> PIDNet.forward -> D2H of the logits
        e2e_steps = max(4, min(a.steps, 12))
        line['e2e'] = run_e2e(model, a, dev, world, e2e_steps, x_host)
        # the same loop on camera frames: 4x fewer bytes up, label maps instead of logits down
        line['e2e_u8_pipeline'] = run_e2e(model, a, dev, world, e2e_steps, x_host, u8=True)
        # ---- batch-1 latency (north-star target < 2 ms), L2 flushed between iterations
        if rank == 0:
            line['latency_bs1'] = run_latency(a, dev, use_graph)
    if cpu_base is not None:
        line['cpu_baseline'] = cpu_base
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)


def run_e2e(model, a, dev, world, steps, x_host, u8=False):
    """Public-API loop: every step copies its pinned host batch to the GPU, calls model(x) and reads the
    logits back; copies are double-buffered against compute on separate streams.
    u8=True: the tools/custom.py pipeline instead -- uint8 BGR frames up, `PIDNet.segment` (fused input transform,
    network, x8 upsample + argmax), uint8 label maps down."""
    import torch.distributed as dist
    B, H, W = a.batch, a.height, a.width
    if u8:
        g = torch.Generator(device='cpu').manual_seed(99)
        f0 = torch.randint(0, 256, (B, H, W, 3), generator=g, dtype=torch.uint8).pin_memory()
        hin = [f0, f0.clone().pin_memory()]
        din = [torch.empty_like(f0, device=dev) for _ in range(2)]
        hout = [torch.empty(B, H, W, dtype=torch.uint8).pin_memory() for _ in range(2)]
    else:
        hin = [x_host, x_host.clone().pin_memory()]
        din = [torch.empty_like(x_host, device=dev) for _ in range(2)]
        hout = [torch.empty(B, a.classes, H // 8, W // 8).pin_memory() for _ in range(2)]
    s_in, s_cmp, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]
    ev_read = [torch.cuda.Event() for _ in range(2)]
    model.use_graph = False

    def loop(n):
        outs = [None, None]
        for i in range(n):
            k = i % 2
            with torch.cuda.stream(s_in):
                if i >= 2:
                    s_in.wait_event(ev_free[k])
                din[k].copy_(hin[k], non_blocking=True)
                ev_in[k].record(s_in)
            with torch.cuda.stream(s_cmp):
                s_cmp.wait_event(ev_in[k])
                if i >= 2:
                    s_cmp.wait_event(ev_read[k])       # previous logits of this slot are on the host
                with torch.no_grad():
                    outs[k] = model.segment(din[k]) if u8 else model(din[k])
                ev_free[k].record(s_cmp)
                ev_done[k].record(s_cmp)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_done[k])
                hout[k].copy_(outs[k], non_blocking=True)
                ev_read[k].record(s_out)
        torch.cuda.synchronize()

    loop(3)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s_in)
    loop(steps)
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = e0.elapsed_time(e1)  # device-timed; the host wall clock is kept beside it
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if u8:
        return {'value': world * B * steps / (ms / 1e3), 'unit': UNIT, 'steps': steps, 'ms_per_step': ms / steps,
                'wall_ms_per_step': wall * 1e3 / steps, 'h2d_bytes_per_step': B * 3 * H * W, 'd2h_bytes_per_step': B * H * W,
                'api': 'pidnet_b200.PIDNet.segment (uint8 HWC BGR frames in, uint8 label maps out: input transform, '
                       'network, x8 upsample + argmax on the device -- the tools/custom.py pipeline)'}
    return {'value': world * B * steps / (ms / 1e3), 'unit': UNIT, 'steps': steps, 'ms_per_step': ms / steps,
            'wall_ms_per_step': wall * 1e3 / steps,
            'h2d_bytes_per_step': B * 3 * H * W * 4, 'd2h_bytes_per_step': B * a.classes * (H // 8) * (W // 8) * 4,
            'api': 'pidnet_b200.PIDNet.forward (fp32 NCHW in, fp32 logits out), H2D/D2H double-buffered on side streams'}


def run_latency(a, dev, use_graph):
    model = make_model(a, dev)
    x = torch.randn(1, 3, a.height, a.width, generator=torch.Generator().manual_seed(0)).to(dev)
    out = torch.empty(1, a.classes, a.height // 8, a.width // 8, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    ts, ts_warm = [], []
    with torch.no_grad():
        for _ in range(10):
            model.forward_into(x, out, use_graph=use_graph)
        for _ in range(30):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            model.forward_into(x, out, use_graph=use_graph)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        # back-to-back (reference harness protocol, models/speed/pidnet_speed.py:238-271: no flush)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            model.forward_into(x, out, use_graph=use_graph)
        e1.record()
        torch.cuda.synchronize()
        b2b = e0.elapsed_time(e1) / 200
    med = statistics.median(ts)
    return {'ms_median_l2_flushed': med, 'ms_min_l2_flushed': min(ts), 'ms_back_to_back': b2b, 'fps_back_to_back': 1e3 / b2b,
            'launches': model.num_launches(), 'published_rtx3090_fp32_fps': 93.2,
            'vs_published_rtx3090': (1e3 / b2b) / 93.2}


def main():
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)


if __name__ == '__main__':
    main()
