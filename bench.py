#!/usr/bin/env python
"""Benchmark of the hot path: grid_pull + grid_push, 256^3, cubic, fp32.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--size 256] [--no-cpu-baseline]

One "step" = one grid_pull of a (1,1,S,S,S) volume through a dense smooth
deformation followed by one grid_push of the pulled image back through the same
deformation (the forward/adjoint pair a registration iteration runs), per GPU.
Metric: Mvoxels/s = (pulled voxels + pushed voxels) / time, whole job.

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definitions
of `value`, `e2e`, `roofline` and `cpu_baseline`.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, 'torch-interpol_b200'))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

ORDER, BOUND, EXTRAPOLATE = 3, 'dct2', True
BOUND_CODE = 3
BYTES_PER_VOXEL = 20   # fp32, D=3, C=1: 12 (grid) + 4 (read) + 4 (write); SURVEY 8(d)


def make_workload(size, device, seed=1234, channels=1, batch=1, dtype=torch.float32, incoherent=False, dim=3):
    """N(0,1) volume; grid = identity + randn(B,D,8,..,8)*3 voxels
    up-sampled (tri)linearly (|disp| <~ 10 voxels, ~5 % of samples out of bounds);
    `incoherent` adds randn * 20 voxels on top (scatter stress, cfg 4).  16-bit
    types are generated in float32 and rounded."""
    g = torch.Generator(device='cpu').manual_seed(seed)
    vol = torch.randn([batch, channels] + [size] * dim, generator=g).to(device)
    coarse = (torch.randn([batch, dim] + [8] * dim, generator=g) * 3.0).to(device)
    mode = {2: 'bilinear', 3: 'trilinear'}[dim]
    disp = torch.nn.functional.interpolate(coarse, size=[size] * dim, mode=mode, align_corners=True)
    ar = torch.arange(size, dtype=torch.float32, device=device)
    ident = torch.stack(torch.meshgrid(*([ar] * dim), indexing='ij'), dim=-1)
    grid = disp.movedim(1, -1) + ident
    if incoherent:
        gd = torch.Generator(device=device).manual_seed(seed + 1)
        grid = grid + torch.randn(grid.shape, generator=gd, device=device) * 20.0
    return vol.to(dtype), grid.contiguous().to(dtype)


class ClockSampler(threading.Thread):
    """SM clock / throttle-reason samples DURING the timed region, through NVML (a poll every ~2 ms;
    the `nvidia-smi -lms` loop of the profiling recipe needs ~1 s to print its first line, longer
    than the whole timed region)."""
    NAMES = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.armed = index, [], False, False
        self.nvml = self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # LOCAL_RANK indexes CUDA_VISIBLE_DEVICES; NVML indexes the board
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = index
            if vis:
                try:
                    phys = int(vis.split(',')[index])
                except Exception:
                    phys = index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def run(self):
        if self.nvml is None:
            return
        n = self.nvml
        masks = [getattr(n, 'nvmlClocksEventReasonHwSlowdown', 0x8), getattr(n, 'nvmlClocksEventReasonHwThermalSlowdown', 0x40),
                 getattr(n, 'nvmlClocksEventReasonSwThermalSlowdown', 0x20), getattr(n, 'nvmlClocksEventReasonSwPowerCap', 0x4)]
        while not self.stop_flag:
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                try:
                    r = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    r = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                if self.armed:
                    self.samples.append((mhz, [bool(r & m) for m in masks]))
            except Exception:
                pass
            time.sleep(0.002)

    def finish(self):
        self.stop_flag = True
        sm, reasons = [], set()
        for mhz, flags in self.samples:
            sm.append(mhz)
            for nm, f in zip(self.NAMES, flags):
                if f:
                    reasons.add(nm)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': getattr(self, 'sm_max', None), 'reasons': [], 'samples': 0}
        return {'sm_mhz': statistics.median(sm), 'sm_max_mhz': self.sm_max, 'reasons': sorted(reasons),
                'samples': len(sm)}


def measured_peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(p) as f:
            return float(json.load(f)['hbm_gbs']), 'measured'
    except Exception:
        return 6650.0, 'fallback'


# ---------------------------------------------------------------------------
# CPU arm: the reference itself when baseline/_ref holds it, else the oracle port
# ---------------------------------------------------------------------------

def load_reference():
    ref_dir = os.path.join(ROOT, 'baseline', '_ref')
    if os.path.isdir(os.path.join(ref_dir, 'interpol')):
        sys.path.insert(0, ref_dir)
        try:
            import warnings
            warnings.filterwarnings('ignore')
            import interpol  # noqa: F401
            return interpol
        except Exception:
            pass
        finally:
            sys.path.remove(ref_dir)
    return None


def cpu_step_fn(size, slab):
    """Returns (fn, kind, cores, sample description, voxels per call).  fn() runs
    pull+push of the first `slab` x-planes of the output lattice on the CPU."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    vol, grid = make_workload(size, 'cpu')
    grid_s = grid[:, :slab].contiguous()
    nvox = 2 * slab * size * size
    sample = 'pull+push of the first %d of %d x-planes of the %d^3 lattice (full %d^3 volume)' % (slab, size, size, size)
    ref = load_reference()
    if ref is not None:
        def fn():
            out = ref.grid_pull(vol, grid_s, interpolation=ORDER, bound=BOUND, extrapolate=EXTRAPOLATE)
            ref.grid_push(out, grid_s, shape=[size] * 3, interpolation=ORDER, bound=BOUND, extrapolate=EXTRAPOLATE)
        return fn, 'reference', torch.get_num_threads(), sample + '; reference TorchScript path', nvox
    import oracle
    oracle.set_num_threads(cores)
    v, g = vol.numpy(), grid_s.numpy()

    def fn():
        out = oracle.grid_pull(v, g, [BOUND_CODE], [ORDER], 1)
        oracle.grid_push(out, g, [size] * 3, [BOUND_CODE], [ORDER], 1, nthreads=cores)
    return fn, 'port', cores, sample + '; C oracle port, OpenMP', nvox


def time_cpu(size, budget_s, steps=1, warmup=1):
    """Times the CPU arm on a slab sized for ~budget_s seconds of work."""
    slab = max(1, size // 32)
    fn, kind, cores, sample, nvox = cpu_step_fn(size, slab)
    fn()                                    # warm-up (TorchScript compile / page-in)
    t0 = time.perf_counter(); fn(); t1 = time.perf_counter() - t0
    per_plane = t1 / slab
    want = int(max(1, min(size, budget_s / max(per_plane, 1e-9) / max(steps + warmup, 1))))
    if want != slab:
        slab = want
        fn, kind, cores, sample, nvox = cpu_step_fn(size, slab)
    for _ in range(warmup):
        fn()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter(); fn(); times.append(time.perf_counter() - t0)
    t = sum(times) / len(times)
    return {'value': nvox / t / 1e6, 'unit': 'Mvoxels/s', 'cores': cores, 'kind': kind,
            'sample': sample, 'ms_per_step': t * 1e3}, t


def run_reference_arm(args, rank):
    if rank != 0:
        return
    cb, t = time_cpu(args.size, budget_s=120.0, steps=args.steps, warmup=min(args.warmup, 1))
    line = {
        'impl': 'reference', 'metric': 'Mvoxels/s grid_pull+grid_push 256^3 cubic fp32', 'value': cb['value'],
        'unit': 'Mvoxels/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': cb['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args.size),
        'cpu_baseline': {k: cb[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
        'e2e': {'value': cb['value'], 'unit': 'Mvoxels/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def workload_config(size):
    return {'workload': '3D %d^3 fp32 C=1, smooth deformation (identity + trilinear-upsampled randn(8^3)*3), '
                        'order=3 cubic, bound=dct2, extrapolate=True, grid_pull then grid_push; one volume per GPU'
                        % size,
            'size': size, 'order': ORDER, 'bound': BOUND, 'channels': 1,
            'cache': 'inputs larger than L2 (volume+grid+output = %.0f MB per op vs 126 MB L2)'
                     % (size ** 3 * BYTES_PER_VOXEL / 1e6)}


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--size', type=int, default=256)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true', help='skip the host-buffer leg (profiling runs)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))

    if args.impl == 'reference':
        run_reference_arm(args, rank)
        return

    import interpol_b200 as ib
    from interpol_b200 import pushpull as pp

    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the product path has no CPU fallback); '
                         'use --impl reference for the CPU arm')
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=device)

    size = args.size
    nvox = size ** 3
    vol, grid = make_workload(size, device, seed=1234 + rank)     # batch sharding: one volume per rank
    bound, order = [BOUND_CODE], [ORDER]

    def step():
        out = pp.grid_pull(vol, grid, bound, order, 1)
        back = pp.grid_push(out, grid, [size] * 3, bound, order, 1)
        return out, back

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    launches0 = ib.launch_count()
    barrier()
    t_start = torch.cuda.Event(enable_timing=True); t_end = torch.cuda.Event(enable_timing=True)
    sampler.armed = True
    t_start.record()
    for k in range(args.steps):
        ev[k][0].record()
        out = pp.grid_pull(vol, grid, bound, order, 1)
        ev[k][1].record()
        back = pp.grid_push(out, grid, [size] * 3, bound, order, 1)
        ev[k][2].record()
    t_end.record()
    barrier()
    sampler.armed = False
    launches = ib.launch_count() - launches0
    clocks = sampler.finish() if rank == 0 else None
    total_ms = t_start.elapsed_time(t_end)
    pull_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / args.steps
    push_ms = sum(e[1].elapsed_time(e[2]) for e in ev) / args.steps
    t = torch.tensor([total_ms], dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * 2 * nvox / (ms_per_step * 1e-3) / 1e6

    # ---- end-to-end through the public API with pinned host buffers ----------
    vol_h = vol.cpu().pin_memory(); grid_h = grid.cpu().pin_memory()
    e2e_steps = max(3, min(args.steps, 5)) if not args.no_e2e else 1

    def e2e_step():
        o = ib.grid_pull(vol_h, grid_h, interpolation=ORDER, bound=BOUND, extrapolate=EXTRAPOLATE)
        b = ib.grid_push(o, grid_h, interpolation=ORDER, bound=BOUND, extrapolate=EXTRAPOLATE)
        return o, b
    for _ in range(3 if not args.no_e2e else 1):      # warm-up: page-locked result buffers reach their steady state
        o_h, b_h = e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        o_h, b_h = e2e_step()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    t = torch.tensor([e2e_s], dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    h2d = (vol_h.numel() + grid_h.numel()) * 4 + (o_h.numel() + grid_h.numel()) * 4
    d2h = (o_h.numel() + b_h.numel()) * 4
    e2e = {'value': world * 2 * nvox / e2e_s / 1e6, 'unit': 'Mvoxels/s', 'h2d_bytes_per_step': h2d,
           'd2h_bytes_per_step': d2h, 'ms_per_step': e2e_s * 1e3, 'steps': e2e_steps,
           'api': 'interpol_b200.grid_pull / grid_push on pinned CPU tensors'}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ------------------------------------
    peak, peak_kind = measured_peak()
    dom = 'push' if push_ms >= pull_ms else 'pull'
    dom_ms = max(push_ms, pull_ms)
    achieved = nvox * BYTES_PER_VOXEL / (dom_ms * 1e-3) / 1e9
    roofline = {'bound': 'hbm', 'kernel': 'grid_%s' % dom, 'achieved': achieved, 'peak': peak, 'peak_kind': peak_kind,
                'unit': 'GB/s', 'frac': achieved / peak, 'traffic': None,
                'algorithmic_bytes_per_launch': nvox * BYTES_PER_VOXEL,
                'pull_ms': pull_ms, 'push_ms': push_ms,
                'pull_frac': nvox * BYTES_PER_VOXEL / (pull_ms * 1e-3) / 1e9 / peak,
                'push_frac': nvox * BYTES_PER_VOXEL / (push_ms * 1e-3) / 1e9 / peak,
                'pull_read_only_frac': nvox * 16 / (pull_ms * 1e-3) / 1e9 / peak}
    prof = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(prof):
        try:
            with open(prof) as f:
                roofline['traffic'] = json.load(f).get('grid_%s' % dom)
        except Exception:
            pass

    line = {
        'metric': 'Mvoxels/s grid_pull+grid_push 256^3 cubic fp32', 'value': value, 'unit': 'Mvoxels/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(size), 'clocks': clocks, 'e2e': e2e, 'gpu_launches': int(launches),
        'roofline': roofline,
    }
    if world == 1 and not args.no_cpu_baseline:
        cb, _ = time_cpu(size, budget_s=20.0, steps=1, warmup=0)
        line['cpu_baseline'] = {k: cb[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
