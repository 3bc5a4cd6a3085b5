#!/usr/bin/env python
"""Benchmark of the hot path.  Default: grid_pull + grid_push, 256^3, cubic, fp32 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--config headline|cfg1|cfg2|cfg3|cfg4|cfg4i|cfg5|backward] [--no-cpu-baseline] [--no-e2e]

One "step" = the ops of the configuration, once, on one batch of synthetic input per GPU.  The default
("headline") is one grid_pull of a (1,1,256,256,256) volume through a dense smooth deformation followed by one
grid_push of the pulled image back through the same deformation (the forward / adjoint pair a registration
iteration runs).  Metric: Mvoxels/s = lattice points processed by every op of the step / time, whole job.
The other configurations are BASELINE.json's configs: cfg1 2-D linear identity pull (latency), cfg2
128^3, cfg3 256^3 C=4 prefilter + pull + grad, cfg4 / cfg4i 256^3 fp16 order-5 dft push + count (smooth /
incoherent grid), cfg5 batch 64 x 192^3 with mixed bounds sharded over the ranks; `backward` is GridPull.backward
at the headline shape.

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definitions of `value`, `e2e`, `roofline`,
`cpu_baseline` and `parity_rel`.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
os.environ.setdefault('IB200_PINNED_POOL_MB', '8192')     # the e2e leg of cfg3 / cfg5 brings GB-sized results back
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
        self.masks = masks
        while not self.stop_flag:
            self.sample_now()
            time.sleep(0.002)

    def sample_now(self):
        """One NVML query, kept if the timed region is open.  Also called by the main thread right after it has
        queued the timed steps (the GPU is then busy with them): an NVML call can take several ms, and a short
        timed region could otherwise end before the polling thread completes a single query."""
        n = self.nvml
        if n is None:
            return
        try:
            armed = self.armed
            mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
            try:
                r = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
            except Exception:
                r = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
            if armed:
                self.samples.append((mhz, [bool(r & m) for m in getattr(self, 'masks', [0x8, 0x40, 0x20, 0x4])]))
        except Exception:
            pass

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
# configurations.  `bytes` = algorithmic bytes per lattice point of each op.
# ---------------------------------------------------------------------------

BOUND_CODES = {'zero': 0, 'replicate': 1, 'dct1': 2, 'dct2': 3, 'dst1': 4, 'dst2': 5, 'dft': 6}


def _spec(name):
    base = dict(name=name, dim=3, size=256, batch=1, channels=1, dtype='f32', order=3, bound=['dct2'],
                extrapolate=True, grid='smooth', ops=['pull', 'push'], sharded=False)
    table = {
        'headline': dict(),
        'cfg1': dict(dim=2, order=1, bound=['zero'], extrapolate=False, grid='identity', ops=['pull']),
        'cfg2': dict(size=128),
        'cfg3': dict(channels=4, ops=['coeff', 'pull', 'grad']),
        'cfg4': dict(dtype='f16', order=5, bound=['dft'], ops=['push', 'count']),
        'cfg4i': dict(dtype='f16', order=5, bound=['dft'], ops=['push', 'count'], grid='incoherent'),
        'cfg5': dict(size=192, batch=64, bound=['dct2', 'dft', 'zero'], sharded=True),
        'backward': dict(ops=['pull_backward']),
    }
    base.update(table[name])
    return base


def _esize(spec):
    return {'f32': 4, 'f16': 2, 'f64': 8}[spec['dtype']]


def op_units_bytes(spec, op, batch):
    """(lattice points, algorithmic bytes) of one op over `batch` volumes (every array touched once)."""
    n = spec['size'] ** spec['dim']
    c, s, d = spec['channels'], _esize(spec), spec['dim']
    per = {'pull': d * s + 2 * c * s, 'push': d * s + 2 * c * s, 'count': d * s + s, 'grad': d * s + c * s + d * c * s,
           'coeff': 2 * c * s, 'pull_backward': (d * s + 2 * c * s) + (d * s + 2 * c * s + d * s)}[op]
    units = n * (c if op == 'coeff' else 1)
    if op == 'coeff':
        per = 2 * s
    return batch * units, batch * units * per


def workload_config(spec, world=1):
    g = {'identity': 'identity grid', 'smooth': 'smooth deformation (identity + (tri)linearly upsampled randn(8^d)*3)',
         'incoherent': 'identity + upsampled randn(8^d)*3 + randn*20 (scatter stress)'}[spec['grid']]
    n = spec['size'] ** spec['dim'] * spec['batch']
    return {'workload': '%s: %dD %s^%d %s C=%d batch=%d, %s, order=%d, bound=%s, extrapolate=%s, ops=%s; %s'
                        % (spec['name'], spec['dim'], spec['size'], spec['dim'], spec['dtype'], spec['channels'], spec['batch'], g,
                           spec['order'], '/'.join(spec['bound']), spec['extrapolate'], '+'.join(spec['ops']),
                           'batch sharded over the ranks' if spec['sharded'] else 'one batch per GPU'),
            'size': spec['size'], 'order': spec['order'], 'bound': spec['bound'][0] if len(spec['bound']) == 1 else spec['bound'],
            'channels': spec['channels'], 'batch': spec['batch'],
            'cache': 'inputs larger than L2 (%.0f MB per op vs 126 MB L2)' % (n * 20 / 1e6) if n * 20 > 2e8
                     else 'L2 flushed between steps by a 256 MB write'}


def make_inputs(spec, device, seed, batch):
    dt = {'f32': torch.float32, 'f16': torch.float16, 'f64': torch.float64}[spec['dtype']]
    vol, grid = make_workload(spec['size'], device, seed=seed, channels=spec['channels'], batch=batch, dtype=dt,
                              incoherent=spec['grid'] == 'incoherent', dim=spec['dim'])
    if spec['grid'] == 'identity':
        ar = torch.arange(spec['size'], dtype=torch.float32, device=device)
        ident = torch.stack(torch.meshgrid(*([ar] * spec['dim']), indexing='ij'), dim=-1)
        grid = ident[None].expand(batch, *ident.shape).contiguous().to(dt)
    return vol, grid


def device_ops(spec, vol, grid):
    """op name -> callable on device tensors through the C-ABI binding layer (interpol_b200.pushpull / coeff)"""
    import interpol_b200 as ib
    from interpol_b200 import pushpull as pp
    bound = [BOUND_CODES[b] for b in spec['bound']]
    order = [spec['order']]
    ex = 1 if spec['extrapolate'] else 0
    shape = list(vol.shape[2:])
    state = {}

    def pull():
        state['pulled'] = pp.grid_pull(state.get('coeff', vol), grid, bound, order, ex)
        return state['pulled']

    def push():
        return pp.grid_push(state.get('pulled', vol), grid, shape, bound, order, ex)

    def pull_backward():
        if 'gout' not in state:
            state['gout'] = torch.randn(vol.shape, generator=torch.Generator(device=vol.device).manual_seed(7),
                                        device=vol.device, dtype=vol.dtype)
            state['vr'] = vol.clone().requires_grad_(); state['gr'] = grid.clone().requires_grad_()
        return pp.grid_pull_backward(state['gout'], state['vr'], state['gr'], bound, order, ex)

    return {
        'pull': pull, 'push': push, 'pull_backward': pull_backward,
        'count': lambda: pp.grid_count(grid, shape, bound, order, ex),
        'grad': lambda: pp.grid_grad(state.get('coeff', vol), grid, bound, order, ex),
        'coeff': lambda: state.__setitem__('coeff', ib.spline_coeff_nd(vol, interpolation=spec['order'], bound=spec['bound'][0],
                                                                      dim=spec['dim'])) or state['coeff'],
    }


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


def cpu_step_fn(spec, slab, inputs=None):
    """Returns (fn, kind, cores, sample description, lattice points per call).  fn() runs the ops of the
    configuration on the CPU for the first `slab` x-planes of the lattice of volume 0 (coeff: channel 0 of
    volume 0) and returns {op: output}.  `inputs`: (volume 0, grid 0) of the GPU arm, so that the timed CPU
    output is the parity oracle of the same run; generated from the configuration's seed otherwise."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    vol, grid = inputs if inputs is not None else make_inputs(spec, 'cpu', 1234, 1)
    size, dim = spec['size'], spec['dim']
    slab = min(slab, size)
    grid_s = grid[:, :slab].contiguous()
    plane = size ** (dim - 1)
    units = {'pull': slab * plane, 'grad': slab * plane, 'push': slab * plane, 'count': slab * plane,
             'coeff': size ** dim, 'pull_backward': slab * plane}
    nvox = sum(units[o] for o in spec['ops'])
    sample = ('%s of the first %d of %d x-planes of the %d^%d lattice of volume 0 (full volume%s)'
              % ('+'.join(spec['ops']), slab, size, size, dim, '; coeff: channel 0' if 'coeff' in spec['ops'] else ''))
    ref = load_reference()
    shape = [size] * dim
    kw = dict(interpolation=spec['order'], bound=spec['bound'] if len(spec['bound']) > 1 else spec['bound'][0],
              extrapolate=spec['extrapolate'])
    # the reference's CPU path has no float16 scatter_add_ / floor: 16-bit configs run it in float32 on the rounded inputs
    cvol, cgrid = (vol.float(), grid_s.float()) if spec['dtype'] == 'f16' else (vol, grid_s)
    if ref is not None:
        def fn():
            out, st = {}, {}
            for op in spec['ops']:
                if op == 'coeff':
                    st['coeff'] = ref.spline_coeff_nd(cvol[:, :1], interpolation=spec['order'], bound=spec['bound'][0], dim=dim)
                    out[op] = st['coeff']
                elif op == 'pull':
                    src = cvol if 'coeff' not in st else torch.cat([st['coeff'], cvol[:, 1:]], 1)
                    st['pulled'] = out[op] = ref.grid_pull(src, cgrid, **kw)
                elif op == 'grad':
                    src = cvol if 'coeff' not in st else torch.cat([st['coeff'], cvol[:, 1:]], 1)
                    out[op] = ref.grid_grad(src, cgrid, **kw)
                elif op == 'push':
                    out[op] = ref.grid_push(st.get('pulled', cvol[:, :, :slab]), cgrid, shape=shape, **kw)
                elif op == 'count':
                    out[op] = ref.grid_count(cgrid, shape=shape, **kw)
                elif op == 'pull_backward':
                    v = cvol.clone().requires_grad_(); g = cgrid.clone().requires_grad_()
                    o = ref.grid_pull(v, g, **kw)
                    o.backward(torch.ones_like(o))
                    out[op] = g.grad
                    out['pull_backward_input'] = v.grad
            return out
        return fn, 'reference', torch.get_num_threads(), sample + '; reference TorchScript path', nvox
    import oracle
    oracle.set_num_threads(cores)
    v, g = cvol.numpy(), cgrid.numpy()
    b = [BOUND_CODES[x] for x in spec['bound']]
    o = [spec['order']]
    ex = 1 if spec['extrapolate'] else 0

    def fn():
        out, st = {}, {}
        for op in spec['ops']:
            if op == 'coeff':
                st['coeff'] = out[op] = oracle.spline_coeff_nd(v[:, :1], b[:1], o, dim)
            elif op == 'pull':
                src = v if 'coeff' not in st else np.concatenate([st['coeff'], v[:, 1:]], 1)
                st['pulled'] = out[op] = oracle.grid_pull(src, g, b, o, ex)
            elif op == 'grad':
                src = v if 'coeff' not in st else np.concatenate([st['coeff'], v[:, 1:]], 1)
                out[op] = oracle.grid_grad(src, g, b, o, ex)
            elif op == 'push':
                out[op] = oracle.grid_push(st.get('pulled', v[:, :, :slab]), g, shape, b, o, ex, nthreads=cores)
            elif op == 'count':
                out[op] = oracle.grid_count(g, shape, b, o, ex, nthreads=cores)
            elif op == 'pull_backward':
                out[op] = (oracle.grid_grad(v, g, b, o, ex)).sum(1)
        return out
    return fn, 'port', cores, sample + '; C oracle port, OpenMP', nvox


def time_cpu(spec, budget_s, steps=1, warmup=1, inputs=None):
    """Times the CPU arm on a slab sized for ~budget_s seconds of work; returns (record, seconds, fn, slab)."""
    size = spec['size']
    slab = max(1, size // 32)
    fn, kind, cores, sample, nvox = cpu_step_fn(spec, slab, inputs)
    fn()                                    # warm-up (TorchScript compile / page-in)
    t0 = time.perf_counter(); fn(); t1 = time.perf_counter() - t0
    per_plane = t1 / slab
    want = int(max(1, min(size, budget_s / max(per_plane, 1e-9) / max(steps + warmup, 1))))
    if want != slab:
        slab = want
        fn, kind, cores, sample, nvox = cpu_step_fn(spec, slab, inputs)
    for _ in range(warmup):
        fn()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter(); fn(); times.append(time.perf_counter() - t0)
    t = sum(times) / len(times)
    return {'value': nvox / t / 1e6, 'unit': 'Mvoxels/s', 'cores': cores, 'kind': kind,
            'sample': sample, 'ms_per_step': t * 1e3}, t, fn, slab


def metric_name(spec):
    if spec['name'] == 'headline':
        return 'Mvoxels/s grid_pull+grid_push 256^3 cubic fp32'
    return 'Mvoxels/s %s %s' % ('+'.join('grid_' + o if o not in ('coeff', 'pull_backward') else
                                         ('spline_coeff_nd' if o == 'coeff' else 'GridPull.backward') for o in spec['ops']), spec['name'])


def run_reference_arm(args, spec, rank):
    if rank != 0:
        return
    cb, t, _, _ = time_cpu(spec, budget_s=120.0, steps=args.steps, warmup=min(args.warmup, 1))
    line = {
        'impl': 'reference', 'metric': metric_name(spec), 'value': cb['value'],
        'unit': 'Mvoxels/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': cb['ms_per_step'], 'higher_is_better': True, 'scaling': 'strong' if spec['sharded'] else 'weak',
        'vs_baseline': None, 'dtype': spec['dtype'], 'data': 'synthetic',
        'config': workload_config(spec),
        'cpu_baseline': {k: cb[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
        'e2e': {'value': cb['value'], 'unit': 'Mvoxels/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def bind_to_gpu_numa(local_rank):
    """Pin this rank's host threads (and thereby its first-touch page-locked buffers) to the CPUs NVML reports as
    local to its GPU, so that 8 ranks do not all stage through one NUMA node."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get('CUDA_VISIBLE_DEVICES')
        phys = int(vis.split(',')[local_rank]) if vis else local_rank
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return len(allowed)
    except Exception:
        pass
    return None


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='headline',
                    choices=['headline', 'cfg1', 'cfg2', 'cfg3', 'cfg4', 'cfg4i', 'cfg5', 'backward'])
    ap.add_argument('--size', type=int, default=None, help='override the edge length of the configuration')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true', help='skip the host-buffer leg (profiling runs)')
    ap.add_argument('--collectives', action='store_true', help='cfg5: also time gather_batch and push_to_shared')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    spec = _spec(args.config)
    if args.size:
        spec['size'] = args.size

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))

    if args.impl == 'reference':
        run_reference_arm(args, spec, rank)
        return

    import interpol_b200 as ib
    from interpol_b200 import distributed as ibd

    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the product path has no CPU fallback); '
                         'use --impl reference for the CPU arm')
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    numa_cpus = bind_to_gpu_numa(local_rank) if world > 1 else None
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=device)

    # batch of this rank: the whole batch of the configuration per GPU (weak scaling), or this rank's contiguous
    # slice of it (cfg5: batch elements are independent -- no collective on the data path)
    if spec['sharded']:
        lo, hi = ibd.shard_bounds(spec['batch'], world, rank)
        batch_local = hi - lo
        seed = 1234 + lo
    else:
        batch_local = spec['batch']
        seed = 1234 + rank
    vol, grid = make_inputs(spec, device, seed, batch_local)
    ops = device_ops(spec, vol, grid)
    names = spec['ops']
    small = sum(op_units_bytes(spec, o, batch_local)[1] for o in names) < 2e8
    flush = torch.empty(64 << 20, dtype=torch.float32, device=device) if small else None

    def step():
        out = None
        for o in names:
            out = ops[o]()
        return out

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
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)] for _ in range(args.steps)]
    launches0 = ib.launch_count()
    barrier()
    sampler.armed = True
    for k in range(args.steps):
        if flush is not None:
            flush.zero_()                   # L2 flush between steps of a small configuration (outside the events)
        ev[k][0].record()
        for i, o in enumerate(names):
            ops[o]()
            ev[k][i + 1].record()
    if rank == 0:
        sampler.sample_now()                # the timed steps are queued and running
    barrier()
    sampler.armed = False
    launches = ib.launch_count() - launches0
    clocks = sampler.finish() if rank == 0 else None
    op_ms = {o: sum(e[i].elapsed_time(e[i + 1]) for e in ev) / args.steps for i, o in enumerate(names)}
    step_ms = sum(e[0].elapsed_time(e[-1]) for e in ev) / args.steps
    t = torch.tensor([step_ms], dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item())
    units_local = sum(op_units_bytes(spec, o, batch_local)[0] for o in names)
    total_units = units_local * world if not spec['sharded'] else sum(op_units_bytes(spec, o, spec['batch'])[0] for o in names)
    value = total_units / (ms_per_step * 1e-3) / 1e6

    # ---- optional collectives of the sharded configuration (timed apart from the data path) ----
    coll = None
    if args.collectives and dist is not None and spec['sharded']:
        from interpol_b200 import pushpull as pp
        bound = [BOUND_CODES[b] for b in spec['bound']]
        pulled = ops['pull']()
        shp = list(vol.shape[2:])
        push1 = lambda i, g, s: pp.grid_push(i, g, s, bound, [spec['order']], 1)

        def timed(fn, reps=3):
            """max over ranks of the mean device time of `reps` calls, after one untimed call (the first collective
            of a communicator pays NCCL's lazy channel set-up: 58 ms for a 28 MB all-reduce in r2d)"""
            out = fn()
            barrier()
            a = torch.cuda.Event(enable_timing=True); b_ = torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                out = fn()
            b_.record(); torch.cuda.synchronize()
            t_ = torch.tensor([a.elapsed_time(b_) / reps], dtype=torch.float64, device=device)
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
            return float(t_.item()), out

        t1, full = timed(lambda: ibd.gather_batch(pulled, spec['batch']))
        t2, shared = timed(lambda: ibd.push_to_shared(push1, pulled[:1], grid[:1], shp))
        # the splat alone, to separate the all-reduce from the kernel
        t3, _ = timed(lambda: push1(pulled[:1], grid[:1], shp))
        coll = {'gather_batch_ms': t1, 'gather_batch_bytes': int(full.numel() * full.element_size()),
                'push_to_shared_ms': t2, 'push_to_shared_allreduce_bytes': int(shared.numel() * shared.element_size()),
                'push_alone_ms': t3, 'timing': 'mean of 3 calls after 1 warm-up call, max over ranks'}
        del full, shared

    # ---- end-to-end through the public API with pinned host buffers ----------
    e2e = None
    if not args.no_e2e and args.config != 'backward':
        e2e = run_e2e(spec, vol, grid, args, dist, device, world, total_units)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ------------------------------------
    peak, peak_kind = measured_peak()
    dom = max(names, key=lambda o: op_ms[o])
    per_op = {}
    for o in names:
        u, by = op_units_bytes(spec, o, batch_local)
        per_op[o] = {'ms': op_ms[o], 'mvox_s': u / (op_ms[o] * 1e-3) / 1e6, 'algorithmic_bytes': by,
                     'frac': by / (op_ms[o] * 1e-3) / 1e9 / peak}
    by = per_op[dom]['algorithmic_bytes']
    roofline = {'bound': 'hbm', 'kernel': 'grid_%s' % dom if dom != 'coeff' else 'spline_coeff_nd', 'achieved': by / (op_ms[dom] * 1e-3) / 1e9,
                'peak': peak, 'peak_kind': peak_kind, 'unit': 'GB/s', 'frac': per_op[dom]['frac'], 'traffic': None,
                'algorithmic_bytes_per_launch': by, 'ops': per_op}
    if 'pull' in per_op:
        roofline['pull_ms'] = op_ms['pull']; roofline['pull_frac'] = per_op['pull']['frac']
    if 'push' in per_op:
        roofline['push_ms'] = op_ms['push']; roofline['push_frac'] = per_op['push']['frac']
    prof = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(prof) and spec['name'] == 'headline':
        try:
            with open(prof) as f:
                tj = json.load(f)
            roofline['traffic'] = tj.get('grid_%s' % dom)
            roofline['traffic_source'] = tj.get('source')
        except Exception:
            pass

    line = {
        'metric': metric_name(spec), 'value': value, 'unit': 'Mvoxels/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step,
        'higher_is_better': True, 'scaling': 'strong' if spec['sharded'] else 'weak', 'vs_baseline': None,
        'dtype': spec['dtype'], 'data': 'synthetic',
        'config': workload_config(spec, world), 'clocks': clocks, 'gpu_launches': int(launches),
        'roofline': roofline,
    }
    if e2e is not None:
        line['e2e'] = e2e
    if spec['name'] == 'cfg1':
        line['latency_us'] = ms_per_step * 1e3
    if coll is not None:
        line['collectives'] = coll
    if numa_cpus is not None:
        line['host_cpus_bound'] = numa_cpus
    if spec['sharded']:
        line['roofline']['frac_of_n_gpus_peak'] = (sum(op_units_bytes(spec, o, spec['batch'])[1] for o in names)
                                                   / (ms_per_step * 1e-3) / 1e9 / (peak * world))
    if world == 1 and not args.no_cpu_baseline:
        cb, _, fn, slab = time_cpu(spec, budget_s=20.0, steps=1, warmup=0, inputs=(vol[:1].cpu(), grid[:1].cpu()))
        line['cpu_baseline'] = {k: cb[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}
        try:
            line['parity_rel'], line['parity'] = parity_vs_cpu(spec, vol, grid, fn, slab)
        except Exception as e:               # never lose the line to the cross-check
            line['parity'] = 'failed: %r' % (e,)
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def parity_vs_cpu(spec, vol, grid, cpu_fn, slab):
    """max|gpu - cpu| / max|cpu| per op, the GPU ops run (untimed) on exactly the sample the CPU arm just timed:
    rank 0's volume 0, first `slab` x-planes of the lattice, prefilter on channel 0 only."""
    from interpol_b200 import pushpull as pp
    cpu = cpu_fn()
    v0, g0 = vol[:1], grid[:1, :slab].contiguous()
    got = {}
    src = v0
    if 'coeff' in spec['ops']:
        got['coeff'] = device_ops(spec, v0[:, :1].contiguous(), g0)['coeff']()
        src = torch.cat([got['coeff'], v0[:, 1:]], 1)
    ops = device_ops(spec, src, g0)
    for o in spec['ops']:
        if o in ('pull', 'grad', 'count'):
            got[o] = ops[o]()
        elif o == 'push':
            img = got['pull'] if 'pull' in got else v0[:, :, :slab].contiguous()
            got[o] = pp.grid_push(img, g0, list(v0.shape[2:]), [BOUND_CODES[b] for b in spec['bound']], [spec['order']],
                                  1 if spec['extrapolate'] else 0)
        elif o == 'pull_backward':
            # the CPU arm back-propagates ones through grid_pull: same cotangent here (the timed steps use randn)
            ones = torch.ones([1, v0.shape[1], *g0.shape[1:-1]], device=v0.device, dtype=v0.dtype)
            gi, gg = pp.grid_pull_backward(ones, v0.clone().requires_grad_(), g0.clone().requires_grad_(),
                                           [BOUND_CODES[b] for b in spec['bound']], [spec['order']], 1 if spec['extrapolate'] else 0)
            got[o] = gg
            if 'pull_backward_input' in cpu:
                got['pull_backward_input'] = gi
    worst, per = 0.0, {}
    for o, g in got.items():
        ref = cpu[o]
        ref = ref.detach().double().numpy() if hasattr(ref, 'detach') else np.asarray(ref, dtype=np.float64)
        g = g.detach().double().cpu().numpy().reshape(ref.shape)
        den = np.abs(ref).max()
        per[o] = float(np.abs(g - ref).max() / den) if den > 0 else float(np.abs(g).max())
        worst = max(worst, per[o])
    return worst, per


def run_e2e(spec, vol, grid, args, dist, device, world, total_units):
    """Same metric through the public API (`interpol_b200.grid_pull` ...) on page-locked HOST tensors: every step
    uploads its inputs and brings every result back.  Inside `stage_scope` each distinct host tensor goes up once per
    step (the grid is used by every op of the step) and a result fed to the next op keeps its device twin."""
    import interpol_b200 as ib
    vol_h = vol.cpu().pin_memory(); grid_h = grid.cpu().pin_memory()
    kw = dict(interpolation=spec['order'], bound=spec['bound'] if len(spec['bound']) > 1 else spec['bound'][0],
              extrapolate=spec['extrapolate'])
    shape = list(vol.shape[2:])
    names = spec['ops']

    def e2e_step():
        outs = []
        with ib.stage_scope():
            src = vol_h
            for o in names:
                if o == 'coeff':
                    src = ib.spline_coeff_nd(vol_h, interpolation=spec['order'], bound=spec['bound'][0], dim=spec['dim'])
                    outs.append(src)
                elif o == 'pull':
                    outs.append(ib.grid_pull(src, grid_h, **kw))
                elif o == 'grad':
                    outs.append(ib.grid_grad(src, grid_h, **kw))
                elif o == 'push':
                    img = outs[-1] if 'pull' in names else vol_h
                    outs.append(ib.grid_push(img, grid_h, shape=shape, **kw))
                elif o == 'count':
                    outs.append(ib.grid_count(grid_h, shape=shape, **kw))
        return outs

    steps = max(3, min(args.steps, 5))
    for _ in range(3):                       # warm-up: page-locked result buffers reach their steady state
        outs = e2e_step()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        outs = e2e_step()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / steps
    t = torch.tensor([e2e_s], dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    esz = vol_h.element_size()
    h2d = (vol_h.numel() + grid_h.numel()) * esz if names != ['count'] else grid_h.numel() * esz
    d2h = sum(o.numel() * o.element_size() for o in outs)
    return {'value': total_units / e2e_s / 1e6, 'unit': 'Mvoxels/s', 'h2d_bytes_per_step': int(h2d),
            'd2h_bytes_per_step': int(d2h), 'ms_per_step': e2e_s * 1e3, 'steps': steps,
            'api': 'interpol_b200.%s on pinned CPU tensors inside interpol_b200.stage_scope() (one upload per '
                   'distinct host tensor per step; the lattice of a pull / grad goes up in slabs so that upload, kernel '
                   'and download overlap; every result copied back)' % ' / '.join(
                       'spline_coeff_nd' if o == 'coeff' else 'grid_' + o for o in names)}


if __name__ == '__main__':
    main()
