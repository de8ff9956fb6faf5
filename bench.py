#!/usr/bin/env python
"""bench.py -- PIDNet hot-path benchmark (contract: see DESIGN.md "Measurement").

  python bench.py --gpus N --steps K --warmup W          # our engine; under torchrun for N > 1
  python bench.py --impl reference --steps K --warmup W  # the reference's own CPU forward (baseline/_ref, else the oracle port)

A "step" is one eval forward of a batch of `--batch` synthetic 3x1024x2048 images per GPU through the PIDNet-S engine
(BASELINE.json configs[1]).  `value` = images/s over all ranks with the inputs resident in HBM; `e2e` = the same metric
through the public API (`PIDNet.forward`) with HOST pinned inputs and the logits read back to the host every step
(H2D/D2H inside the timed region, copies double-buffered against compute).  The line also carries
  roofline       the kernel family with the largest share of the step, against the measured burst AND sustained peaks;
  cpu_baseline   the reference's CPU forward on the box's host cores (bounded sample);
  ref_gpu_eager  the reference module itself through torch eager / cuDNN on this GPU (fp32 NCHW as shipped, TF32 on,
                 bf16 channels_last) at batch 1 and at the bench batch -- the existing GPU path, same harness protocol
                 as models/speed/pidnet_speed.py:238-271;
  train          BASELINE.json configs[4]: PIDNet-S fwd + OHEM/boundary loss + bwd + NCCL gradient all-reduce + SGD,
                 12 x 3 x 1024 x 1024 per GPU (at every --gpus N).
One JSON line is printed by rank 0.
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

UNIT = 'img/s'
REF_DIR = os.path.join(ROOT, 'baseline', '_ref')
CITYSCAPES_W = [0.8373, 0.918, 0.866, 1.0345, 1.0166, 0.9969, 0.9754, 1.0489, 0.8786, 1.0023, 0.9539, 0.9843, 1.1116,
                0.9037, 1.0865, 1.0955, 1.0865, 1.1529, 1.0507]      # datasets/cityscapes.py:55-59


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
    ap.add_argument('--skip-train', action='store_true', help='skip the training-step record (config 5)')
    ap.add_argument('--skip-ref-gpu', action='store_true', help='skip the torch-eager/cuDNN run of the reference module')
    ap.add_argument('--profile-out', default=None, help='write the per-launch table (json) here')
    return ap.parse_args()


def metric_name(a):
    return f'{a.model}_{a.height}x{a.width}_images_per_sec'


def workload_config(a, world):
    cfgno = {('pidnet_s', 1024, 2048): 'configs[1]', ('pidnet_m', 720, 960): 'configs[2]', ('pidnet_l', 1024, 2048): 'configs[3]'}
    return {
        'workload': f'{a.model} eval forward, synthetic {a.batch}x3x{a.height}x{a.width} fp32 NCHW per GPU -> '
                    f'fp32 logits [{a.batch},{a.classes},{a.height // 8},{a.width // 8}] (BASELINE.json '
                    f'{cfgno.get((a.model, a.height, a.width), "geometry not in configs")})',
        'batch_per_gpu': a.batch, 'global_batch': a.batch * world, 'height': a.height, 'width': a.width,
        'classes': a.classes, 'weights': 'random init, seed 0', 'parallelism': f'batch-sharded x{world}, no collective',
        'l2': f'per-step input ({a.batch * 3 * a.height * a.width * 4 / 1e6:.0f} MB) and activations exceed the 126 MB L2; '
              'no explicit flush needed for the throughput loop; the bs1 latency loop flushes L2 between iterations',
    }


def cpu_model_string():
    try:
        with open('/proc/cpuinfo') as f:
            for line in f:
                if line.startswith('model name'):
                    return line.split(':', 1)[1].strip()
    except OSError:
        pass
    return ''


# ------------------------------------------------------------------------------------------- the reference itself
def load_reference_models():
    """The UNMODIFIED reference module (models/pidnet.py + models/model_utils.py), vendored by `__graft_entry__.build()` into the
    git-ignored baseline/_ref/ when /root/reference is present; None where it is absent."""
    if not os.path.exists(os.path.join(REF_DIR, 'models', 'pidnet.py')):
        return None
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    try:
        import importlib
        return importlib.import_module('models.pidnet')
    except Exception as exc:   # noqa: BLE001
        sys.stderr.write(f'bench.py: baseline/_ref is not importable ({exc}); falling back to the oracle port\n')
        return None


def reference_forward_fn(a, device):
    """(callable x -> logits, kind): the reference's `get_pred_model(name, classes)` in eval mode on `device` when baseline/_ref
    exists ("reference"), else the oracle's restatement of the same op sequence ("port")."""
    ref = load_reference_models()
    torch.manual_seed(0)
    if ref is not None:
        model = ref.get_pred_model(a.model, a.classes).eval().to(device)
        return model, 'reference', 'baseline/_ref/models/pidnet.py get_pred_model (unmodified reference module)'
    from oracle import pidnet_oracle as O
    cfg = O.config_for(a.model, a.classes, False)
    sd = {k: v.to(device) for k, v in O.make_state_dict(cfg, seed=0, randomize_bn=False).items()}
    return (lambda x: O.pidnet_forward(sd, x)), 'port', 'oracle/pidnet_oracle.py (restatement of the reference op sequence)'


# ------------------------------------------------------------------------------------------- CPU arm
def cpu_reference_forward(a, steps, warmup):
    """The reference's CPU forward for this path on all host threads.  One step = ONE image of the workload (bounded sample)."""
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    fn, kind, what = reference_forward_fn(a, torch.device('cpu'))
    x = torch.randn(1, 3, a.height, a.width, generator=torch.Generator().manual_seed(0))
    ts = []
    with torch.no_grad():
        for _ in range(warmup):
            fn(x)
        for _ in range(steps):
            t = time.perf_counter()
            fn(x)
            ts.append(time.perf_counter() - t)
    return dict(value=1.0 / statistics.median(ts), unit=UNIT, cores=cores, kind=kind,
                sample=f'{steps} forwards of 1x3x{a.height}x{a.width} fp32 (1 image of the batch) after {warmup} warm-ups, median; '
                       f'{what} on torch CPU ({torch.get_num_threads()} threads)',
                cpu=cpu_model_string(), median_s=statistics.median(ts), min_s=min(ts))


def run_reference(a):
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if rank != 0:
        return
    steps = max(1, min(a.steps, 40))
    warmup = max(5, min(a.warmup, 8))
    cb = cpu_reference_forward(a, steps, warmup)
    line = {
        'impl': 'reference', 'metric': metric_name(a), 'value': cb['value'], 'unit': UNIT, 'n_gpus': a.gpus, 'steps': steps,
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


# ------------------------------------------------------------------------------------------- host placement
def bind_to_gpu_numa(index):
    """Pin this process to the CPU cores NVML reports as local to GPU `index` (its NUMA node) BEFORE the pinned host buffers are
    allocated, so that first-touch places them in the memory the GPU's PCIe root complex reaches without crossing sockets.
    Returns a short description for the JSON line (or why nothing was done)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        local = {w * 64 + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        allowed = os.sched_getaffinity(0)
        pick = sorted(local & allowed)
        if not pick:
            return f'gpu-local cores {sorted(local)[:4]}.. not in the allowed set; unchanged ({len(allowed)} cores)'
        if set(pick) == set(allowed):
            return f'all {len(allowed)} allowed cores are gpu-local; unchanged'
        os.sched_setaffinity(0, pick)
        return f'bound to {len(pick)} gpu-local cores ({pick[0]}..{pick[-1]}) of {len(allowed)} allowed'
    except Exception as exc:   # noqa: BLE001
        return f'unavailable ({type(exc).__name__}: {exc})'


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
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed `ncu --set full` capture of this
    same command (profiles/r02/traffic.json, else r01; written by tools/ncu_traffic.py) -- read from the file, not measured in
    this run; None if that kernel was not captured."""
    for rnd in ('r02', 'r01'):
        path = os.path.join(ROOT, 'profiles', rnd, 'traffic.json')
        try:
            with open(path) as f:
                t = json.load(f)
            e = t.get(kernel)
            if e is not None:
                return e['dram_bytes_per_launch'], f'profiles/{rnd}/traffic.json (committed ncu --set full capture, not measured in this run)'
        except (OSError, ValueError, KeyError):
            continue
    return None, None


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
        cpu_base = cpu_reference_forward(a, steps=5, warmup=5)
    placement = bind_to_gpu_numa(local)   # after the all-core CPU baseline, before any pinned allocation

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
        'metric': metric_name(a), 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': a.steps, 'warmup': max(a.warmup, 3),
        'ms_per_step': ms / a.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16',
        'data': 'synthetic', 'config': workload_config(a, world), 'clocks': clocks,
        'gpu_launches': launches * a.steps, 'launches_per_step': launches, 'cuda_graph': use_graph,
        'conv_gflop_per_image': flops_per_img / 1e9, 'host_placement': placement,
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
        traffic, traffic_src = measured_traffic(dom['kernel'])
        # the launches are timed one at a time, each alone on an idle GPU at full clock: the BURST peak is the matching
        # denominator (`frac`); the sustained one (what a long back-to-back run can hold) is quoted beside it
        line['roofline'] = {
            'kernel': dom['kernel'], 'bound': 'tensor' if tensor_bound else 'hbm',
            'achieved': tf if tensor_bound else gbs, 'peak': peaks['tc_burst'] if tensor_bound else peaks['hbm'],
            'unit': 'TFLOP/s' if tensor_bound else 'GB/s',
            'frac': (tf / peaks['tc_burst']) if tensor_bound else (gbs / peaks['hbm']),
            'frac_burst': (tf / peaks['tc_burst']) if tensor_bound else (gbs / peaks['hbm']),
            'frac_sustained': (tf / peaks['tc_sustained']) if tensor_bound else (gbs / peaks['hbm']),
            'peak_burst': peaks['tc_burst'] if tensor_bound else peaks['hbm'],
            'peak_sustained': peaks['tc_sustained'] if tensor_bound else peaks['hbm'],
            'traffic': traffic, 'traffic_source': traffic_src,
            'peak_source': peaks['source'] + (' bf16 burst (kernel timed alone)' if tensor_bound else ' hbm copy'),
            'launches_per_step': dom['launches'], 'avg_launch_ms': dom['ms'] / dom['launches'],
            'algorithmic_gflop_per_launch': dom['flops'] / dom['launches'] / 1e9,
            'algorithmic_mb_per_launch': dom['bytes'] / dom['launches'] / 1e6,
            'share_of_step': dom['ms'] / tot_ms, 'hbm_gbs_same_kernel': gbs, 'hbm_frac_same_kernel': gbs / peaks['hbm'],
        }
        line['kernels'] = [dict(kernel=k['kernel'], launches=k['launches'], ms=round(k['ms'], 4),
                                share=round(k['ms'] / tot_ms, 4),
                                tflops=round(k['flops'] / (k['ms'] * 1e-3) / 1e12, 2),
                                gbs=round(k['bytes'] / (k['ms'] * 1e-3) / 1e9, 1)) for k in top[:10]]
        whole = flops_per_img * B / (ms / a.steps * 1e-3) / 1e12
        line['whole_net'] = {'conv_tflops': whole, 'frac_of_tc_sustained': whole / peaks['tc_sustained'],
                             'frac_of_tc_burst': whole / peaks['tc_burst'], 'serialised_sum_ms': tot_ms}
        if a.profile_out and rank == 0:
            with open(a.profile_out, 'w') as f:
                json.dump(rows, f, indent=1)

        # ---- e2e through the public API: pinned host input -> H2D -> PIDNet.forward -> D2H of the logits
        e2e_steps = max(4, min(a.steps, 12))
        line['e2e'] = run_e2e(model, a, dev, world, e2e_steps, x_host, use_graph)
        # the same loop on camera frames: 4x fewer bytes up, label maps instead of logits down
        line['e2e_u8_pipeline'] = run_e2e(model, a, dev, world, e2e_steps, x_host, use_graph, u8=True)
        # ---- batch-1 latency (north-star target < 2 ms), L2 flushed between iterations
        if rank == 0:
            line['latency_bs1'] = run_latency(a, dev, use_graph)
    del x, out
    torch.cuda.empty_cache()
    if not a.skip_train:
        line['train'] = run_train(a, dev, world, rank, peaks)
    if rank == 0 and world == 1 and not a.skip_ref_gpu and not a.skip_extras:
        try:
            line['ref_gpu_eager'] = run_ref_gpu_eager(a, dev, line)
        except Exception as exc:   # noqa: BLE001  (the competitor's failure must not take the bench line down)
            line['ref_gpu_eager'] = {'unavailable': f'{type(exc).__name__}: {exc}'}
    if cpu_base is not None:
        line['cpu_baseline'] = cpu_base
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)


def measure_h2d_ceiling(dst, src, world, dev):
    """Plain cudaMemcpyAsync of the step's pinned input, one call per copy, all ranks at once: what the host memory / PCIe path
    can deliver per GPU while every GPU of the job is copying (the ceiling of the e2e loop)."""
    import torch.distributed as dist
    for _ in range(2):
        dst.copy_(src, non_blocking=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        dst.copy_(src, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    gbs = 4 * src.numel() * src.element_size() / (e0.elapsed_time(e1) * 1e-3) / 1e9
    if world > 1:
        t = torch.tensor([gbs, gbs], device=dev)
        mn = t[:1].clone()
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        dist.all_reduce(t[1:], op=dist.ReduceOp.SUM)
        return float(mn.item()), float(t[1].item())
    return gbs, gbs


def run_e2e(model, a, dev, world, steps, x_host, use_graph, u8=False):
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
        dlab = [torch.empty(B, H, W, dtype=torch.uint8, device=dev) for _ in range(2)]
        dlog = [torch.empty(B, a.classes, H // 8, W // 8, device=dev) for _ in range(2)]
    else:
        hin = [x_host, x_host.clone().pin_memory()]
        din = [torch.empty_like(x_host, device=dev) for _ in range(2)]
        hout = [torch.empty(B, a.classes, H // 8, W // 8).pin_memory() for _ in range(2)]
    ceil_min, ceil_sum = measure_h2d_ceiling(din[0], hin[0], world, dev)
    s_in, s_cmp, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]
    ev_read = [torch.cuda.Event() for _ in range(2)]
    model.use_graph = use_graph     # the public call replays the CUDA graph captured for this (input, output) pointer set

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
                    outs[k] = model.segment(din[k], out=dlab[k], logits=dlog[k]) if u8 else model(din[k])
                ev_free[k].record(s_cmp)
                ev_done[k].record(s_cmp)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_done[k])
                hout[k].copy_(outs[k], non_blocking=True)
                ev_read[k].record(s_out)
        torch.cuda.synchronize()

    loop(4)
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
    model.use_graph = False
    h2d = B * 3 * H * W * (1 if u8 else 4)
    d2h = B * H * W if u8 else B * a.classes * (H // 8) * (W // 8) * 4
    res = {'value': world * B * steps / (ms / 1e3), 'unit': UNIT, 'steps': steps, 'ms_per_step': ms / steps,
           'wall_ms_per_step': wall * 1e3 / steps, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h, 'cuda_graph': use_graph,
           'h2d_ceiling_gbs_per_gpu_min': ceil_min, 'h2d_ceiling_gbs_all_gpus': ceil_sum,
           'h2d_ceiling_img_per_s': ceil_sum * 1e9 / (h2d / B),
           'h2d_ceiling_note': 'plain cudaMemcpyAsync of the same pinned buffer (one call per copy), all ranks copying at once'}
    if u8:
        res['api'] = ('pidnet_b200.PIDNet.segment (uint8 HWC BGR frames in, uint8 label maps out: input transform, '
                      'network, x8 upsample + argmax on the device -- the tools/custom.py pipeline)')
    else:
        res['api'] = 'pidnet_b200.PIDNet.forward (fp32 NCHW in, fp32 logits out), H2D/D2H double-buffered on side streams'
    return res


def run_latency(a, dev, use_graph):
    model = make_model(a, dev)
    x = torch.randn(1, 3, a.height, a.width, generator=torch.Generator().manual_seed(0)).to(dev)
    out = torch.empty(1, a.classes, a.height // 8, a.width // 8, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    ts = []
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


# ------------------------------------------------------------------------------------------- reference on the GPU (torch eager)
def run_ref_gpu_eager(a, dev, line):
    """The reference module through torch eager / cuDNN on this GPU: the existing GPU path (BASELINE.md B0').
    Harness protocol of models/speed/pidnet_speed.py:238-271 (10 warm-ups, sync before/after), timed with CUDA events."""
    fn, kind, what = reference_forward_fn(a, dev)
    is_module = isinstance(fn, torch.nn.Module)
    out = {'kind': kind, 'what': what + ' on cuda through torch eager (cuDNN / ATen)', 'torch': torch.__version__,
           'cudnn': torch.backends.cudnn.version(), 'runs': []}

    def timed(call, x, iters):
        with torch.no_grad():
            for _ in range(10 if x.shape[0] == 1 else 3):
                call(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                call(x)
            e1.record()
            torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    variants = [('fp32 NCHW, TF32 off (as shipped)', False, None), ('fp32 NCHW, TF32 on', True, None)]
    if is_module:
        variants.append(('bf16 channels_last', True, torch.bfloat16))
    for label, tf32, dt in variants:
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.benchmark = True
        call = fn
        if dt is not None:
            import copy
            m16 = copy.deepcopy(fn).to(dt).to(memory_format=torch.channels_last)
            call = m16
        for bs in (1, a.batch):
            x = torch.randn(bs, 3, a.height, a.width, device=dev)
            if dt is not None:
                x = x.to(dt).contiguous(memory_format=torch.channels_last)
            try:
                ms = timed(call, x, 50 if bs == 1 else 5)
                out['runs'].append({'variant': label, 'batch': bs, 'ms': ms, 'img_per_s': bs * 1e3 / ms})
            except Exception as exc:   # noqa: BLE001
                out['runs'].append({'variant': label, 'batch': bs, 'error': f'{type(exc).__name__}: {str(exc)[:120]}'})
            del x
            torch.cuda.empty_cache()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ours_bs = line['value'] / line['n_gpus']
    ours_b1 = line.get('latency_bs1', {}).get('fps_back_to_back')
    best = lambda bs: max([r['img_per_s'] for r in out['runs'] if r.get('batch') == bs and 'img_per_s' in r], default=None)
    shipped = lambda bs: next((r['img_per_s'] for r in out['runs'] if r.get('batch') == bs and 'as shipped' in r['variant']
                               and 'img_per_s' in r), None)
    out['ours_over_ref'] = {
        f'batch{a.batch}_vs_as_shipped_fp32': ours_bs / shipped(a.batch) if shipped(a.batch) else None,
        f'batch{a.batch}_vs_best_variant': ours_bs / best(a.batch) if best(a.batch) else None,
        'batch1_vs_as_shipped_fp32': ours_b1 / shipped(1) if (ours_b1 and shipped(1)) else None,
        'batch1_vs_best_variant': ours_b1 / best(1) if (ours_b1 and best(1)) else None,
        'note': 'device-resident img/s of the engine divided by the reference module timed the same way (no host copies on either side)',
    }
    return out


# ------------------------------------------------------------------------------------------- training step (config 5)
def run_train(a, dev, world, rank, peaks):
    """BASELINE.json configs[4] / SURVEY 8d config 5: PIDNet-S, augment, train mode, 12 x 3 x 1024 x 1024 fp32 images per GPU,
    labels randint(0,19) with a 255-ignore band, bd_gt = (rand > 0.9), Cityscapes class weights, OHEM 0.9 / 131072,
    BALANCE_WEIGHTS [0.4, 1.0], SB_WEIGHTS 1.0; forward + loss + backward + gradient all-reduce (+ fused SGD step).
    Timed through the reference loop's API (FullModel -> losses.mean().backward() -> optimizer.step(), utils/function.py:43-49)
    and, for the breakdown, through the trainer with the gradient exchange off."""
    import torch.distributed as dist
    from pidnet_b200 import BondaryLoss, FullModel, FusedSGD, OhemCrossEntropy, PIDNet, adjust_learning_rate
    B, H, W = 12, 1024, 1024
    steps, warmup = max(4, min(a.steps, 10)), 3
    torch.manual_seed(0)
    model = PIDNet(m=2, n=3, num_classes=19, planes=32, ppm_planes=96, head_planes=128, augment=True).to(dev)
    weight = torch.tensor(CITYSCAPES_W)
    full = FullModel(model, OhemCrossEntropy(255, 0.9, 131072, weight), BondaryLoss(), return_outputs=False).to(dev).train()
    tr = full.trainer
    opt = FusedSGD(full, lr=0.01, momentum=0.9, weight_decay=5e-4)
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.randn(B, 3, H, W, generator=g).to(dev)
    labels = torch.randint(0, 19, (B, H, W), generator=g)
    labels[:, :32, :] = 255
    labels = labels.to(dev)
    bd = (torch.rand(B, H, W, generator=g) > 0.9).float().to(dev)

    def api_step(i):
        losses, _, _, _ = full(x, labels, bd)
        loss = losses.mean()
        opt.zero_grad()
        loss.backward()                      # network backward; bucketed all-reduce overlapped with it
        opt.step()
        adjust_learning_rate(opt, 0.01, 1000, i)
        return loss

    def local_step(i):                       # same work without the gradient exchange
        tr.step(x, labels, bd, weight, full._crit.cfg, backward=2, want_logits=False)
        tr.backward(x, allreduce=False)
        opt.step()

    def timed(fn):
        for i in range(warmup):
            fn(i)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            r = fn(i)
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, r

    ms_local, _ = timed(local_step)
    ms_api, loss = timed(api_step)           # default exchange: one all-reduce of the flat gradient after the backward
    tr.overlap_allreduce = True              # bucketed all-reduces behind the backward ranges, for comparison
    ms_api_bucketed, _ = timed(api_step) if world > 1 else (ms_api, None)
    tr.overlap_allreduce = False
    import ctypes as C
    f, b = C.c_int(), C.c_int()
    tr.lib.pidnet_train_num_launches(tr.h, C.byref(f), C.byref(b))
    flops = 3 * 50.295e9 * B    # fwd (augment) 50.295 GFLOP / image at 1024x1024, fwd+bwd = 3x
    tfl = flops / (ms_api * 1e-3) / 1e12
    rec = {
        'metric': 'pidnet_s_train_1024x1024_images_per_sec', 'value': world * B * 1e3 / ms_api, 'unit': UNIT, 'n_gpus': world,
        'steps': steps, 'warmup': warmup, 'ms_per_step': ms_api, 'ms_per_step_no_exchange': ms_local,
        'allreduce_ms_exposed': max(0.0, ms_api - ms_local) if world > 1 else 0.0,
        'ms_per_step_bucketed_allreduce': ms_api_bucketed,
        'exchange': f'one NCCL all-reduce of the flat fp32 gradient ({4 * tr.n_param} bytes) after the backward (default); '
                    f'ms_per_step_bucketed_allreduce = {len([r for s in tr.segment_ranges() for r in s])} async all-reduces over '
                    f'{len(tr.segment_ranges())} backward ranges (buckets in reverse layer order, EngineTrainer.overlap_allreduce)'
                    if world > 1 else 'single rank: no exchange',
        'conv_tflops': tfl, 'frac_of_tc_sustained': tfl / peaks['tc_sustained'], 'frac_of_tc_burst': tfl / peaks['tc_burst'],
        'algorithmic_gflop_per_step_per_gpu': flops / 1e9, 'launches': {'forward': f.value, 'backward': b.value},
        'gpu_launches': (f.value + b.value + 12) * steps, 'loss': float(loss.detach()), 'scaling': 'weak',
        'dtype': 'bf16 activations / gradients, fp32 master weights and weight gradients', 'data': 'synthetic',
        'api': 'pidnet_b200.FullModel -> losses.mean().backward() -> FusedSGD.step() (the reference loop, utils/function.py:43-49)',
        'config': {'workload': f'PIDNet-S train fwd + OHEM/boundary loss + bwd + SGD, {B}x3x{H}x{W} per GPU, OHEM 0.9/131072, '
                               'Cityscapes class weights (BASELINE.json configs[4])', 'batch_per_gpu': B,
                   'global_batch': B * world},
    }
    del full, tr, opt, model
    torch.cuda.empty_cache()
    if world == 1 and rank == 0 and not a.skip_ref_gpu:
        try:
            rec['ref_gpu_eager'] = run_ref_gpu_train(dev, x, labels, bd, weight, ms_api)
        except Exception as exc:   # noqa: BLE001
            rec['ref_gpu_eager'] = {'unavailable': f'{type(exc).__name__}: {str(exc)[:160]}'}
        torch.cuda.empty_cache()
    return rec


def run_ref_gpu_train(dev, x, labels, bd, weight, ours_ms):
    """The same training step through torch autograd / cuDNN on this GPU: the reference's PIDNet module (baseline/_ref, train mode,
    augment=True) when vendored, else the oracle's restatement of it; the loss is the oracle's restatement of FullModel /
    OhemCrossEntropy / BondaryLoss (the reference's utils import a yacs config that is absent here) -- it materialises the
    full-resolution logits exactly as the reference does; torch.optim.SGD.  fp32 (TF32 off, as shipped) and bf16 autocast."""
    from oracle import criterion_oracle as CO
    from oracle import pidnet_oracle as O
    ref = load_reference_models()
    out = {'runs': [], 'criterion': 'oracle/criterion_oracle.py (restatement of utils/utils.py:37-57 + utils/criterion.py)',
           'model': 'baseline/_ref models/pidnet.py PIDNet(augment=True).train()' if ref is not None else 'oracle/pidnet_oracle.py (port)'}
    wd = weight.to(dev)
    for label, tf32, autocast in (('fp32, TF32 off (as shipped)', False, False), ('fp32, TF32 on', True, False), ('bf16 autocast', True, True)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.manual_seed(0)
        if ref is not None:
            net = ref.PIDNet(m=2, n=3, num_classes=19, planes=32, ppm_planes=96, head_planes=128, augment=True).to(dev).train()
            params = list(net.parameters())
            fwd = lambda inp: net(inp)
        else:
            cfg = O.config_for('pidnet_s', 19, True)
            sd = {k: v.to(dev) for k, v in O.make_state_dict(cfg, 0, randomize_bn=False).items()}
            params = [v for k, v in sd.items() if v.dtype.is_floating_point and 'running_' not in k]
            for p_ in params:
                p_.requires_grad_(True)
            fwd = lambda inp: O.pidnet_forward(sd, inp, training=True)
        opt = torch.optim.SGD(params, lr=0.01, momentum=0.9, weight_decay=5e-4)

        def step():
            with torch.autocast('cuda', dtype=torch.bfloat16, enabled=autocast):
                outs = fwd(x)
            losses, _, _, _ = CO.full_model_forward([o.float() for o in outs], labels, bd, wd, dict(ohem_keep=131072))
            loss = losses.mean()
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
        try:
            for _ in range(2):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            out['runs'].append({'variant': label, 'ms_per_step': ms, 'img_per_s': x.shape[0] * 1e3 / ms, 'speedup_of_ours': ms / ours_ms})
        except Exception as exc:   # noqa: BLE001
            out['runs'].append({'variant': label, 'error': f'{type(exc).__name__}: {str(exc)[:120]}'})
        del opt, params
        torch.cuda.empty_cache()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return out


def main():
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)


if __name__ == '__main__':
    main()
