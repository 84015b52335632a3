#!/usr/bin/env python
"""bench.py -- FoKL.fit throughput on B200 (BASELINE.json metric: candidate-models/sec; basis-matrix GB/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg4] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one complete `FoKL.fit` (forward selection: basis build -> Gram -> per-candidate spectral
factorisation + Gibbs chain -> BIC) on the synthetic workload of BASELINE.json `configs[3]` (cfg4: N = 10M rows,
8 inputs, 3-way interactions, cubic splines, 1000 + 1000 draws), rows sharded over the N GPUs (strong scaling of the
fixed 10M-row problem, one NCCL allreduce of the new Gram block per substage).  One "candidate model" = one `gibbs`
invocation of the reference (FoKLRoutines.py:1650 / :1681).

    value  candidate-models/sec with the normalised dataset already resident in HBM (fit(DeviceDataset))
    e2e    the same through the public API with HOST numpy buffers: FoKL.fit(inputs, data, clean=True) -- host
           formatting, host->device copies and the device->host read of betas/mtx/evs inside the timed region
    roofline      K1 basis kernel (HBM-bound): algorithmic bytes 8*N*(M + C) per launch / CUDA-event time, against
                  MEASURED_PEAKS.json hbm_gbs; `roofline_gram` the K2 FP64 DMMA Gram against an in-run cuBLAS DGEMM
    cpu_baseline  the CPU oracle (oracle/fokl_oracle.py, a numpy restatement of the reference) on a bounded sample
`--impl reference` times that CPU oracle alone (rank 0), same metric / config keys.
Timing: CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, 'fokl-gpy_b200'), ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

import bench_data  # noqa: E402

METRIC = 'FoKL.fit candidate-models/sec'
UNIT = 'candidate-models/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='cfg4', choices=sorted(bench_data.CONFIGS))
    ap.add_argument('--n', type=int, default=0, help='override the number of rows (debugging only)')
    ap.add_argument('--draws', type=int, default=1000)
    ap.add_argument('--cpu-rows', type=int, default=4000, help='rows of the bounded CPU sample')
    ap.add_argument('--cpu-seconds', type=float, default=25.0, help='wall-time budget of the bounded CPU sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    return ap.parse_args()


def config_dict(a, n_total, world, extra=None):
    c = bench_data.CONFIGS[a.workload]
    d = {'workload': '%s: synthetic N=%d, M=%d, %s, %s, burnin+draws=%d+%d' % (
        a.workload, n_total, c['m'], '3-way' if c['way3'] else '2-way', c['kernel'], a.draws, a.draws),
        'rows': n_total, 'inputs': c['m'], 'way3': c['way3'], 'kernel': c['kernel'],
        'sharding': 'rows/%d' % world, 'l2': 'inputs_exceed_l2 (%.0f MB per rank)' % (n_total / world * c['m'] * 8 / 1e6)}
    if extra:
        d.update(extra)
    return d


# ---------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ---------------------------------------------------------------------------------------------------
class Clocks:
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.th = index, [], threading.Event(), None

    def _run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=10).stdout
                parts = [s.strip() for s in out.strip().split(',')]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            self.stop_flag.wait(0.5)

    def start(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def stop(self):
        self.stop_flag.set()
        if self.th is not None:
            self.th.join(timeout=15)
        sm, mx, reasons = [], 0.0, set()
        for p in self.samples:
            try:
                sm.append(float(p[0]))
                mx = max(mx, float(p[1]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), p[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx or None, 'reasons': sorted(reasons),
                'samples': len(sm)}


# ---------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on a bounded sample
# ---------------------------------------------------------------------------------------------------
class _CpuBudget(Exception):
    pass


def cpu_fit_once(a, phis, rows, budget_s):
    """The CPU oracle's fit on `rows` rows of the workload, cut off after `budget_s` seconds of wall time: returns
    (gibbs calls completed, seconds up to the last completed call, threads, largest model width reached)."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import fokl_oracle as fo
    c = bench_data.CONFIGS[a.workload]
    x, y = bench_data.make_rows(a.workload, 0, rows, n_total=rows)
    np.random.seed(c['seed'])
    threads = os.cpu_count() or 1
    state = dict(calls=0, t=0.0, pmax=0)
    t0 = time.perf_counter()

    def on_gibbs(info):
        state['calls'] = info['call']
        state['t'] = time.perf_counter() - t0
        state['pmax'] = max(state['pmax'], info['discmtx'].shape[0] + 1)
        if budget_s and state['t'] > budget_s:
            raise _CpuBudget()

    try:
        fo.fit(x, y, phis, kernel=c['kernel'], way3=c['way3'], draws=a.draws, burnin=a.draws, threads=threads,
               on_gibbs=on_gibbs)
    except _CpuBudget:
        pass
    return state['calls'], state['t'], threads, state['pmax']


def cpu_sample_desc(a, rows, budget_s, calls, pmax):
    return ('oracle port (numpy/OpenBLAS/LAPACK like the reference + a C basis helper, i.e. faster than the reference, '
            'whose basis build is a Python triple loop) of the %s fit at N=%d rows, %d+%d draws, cut off after %.0f s: '
            'the first %d gibbs calls (model width up to %d columns). The complete CPU fit at N=4000 takes 300 s '
            '(0.62 candidate-models/s on 8 cores, see BASELINE.md)'
            % (a.workload, rows, a.draws, a.draws, budget_s, calls, pmax))


def run_reference(a):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    phis = load_phis(a)
    n_total = a.n or bench_data.CONFIGS[a.workload]['n']
    for _ in range(a.warmup):
        cpu_fit_once(a, phis, a.cpu_rows, 2.0)
    tot_models, tot_s, threads, pmax = 0, 0.0, 1, 0
    budget = a.cpu_seconds * 2.0       # each step: a bounded sample of the fit
    for _ in range(a.steps):
        m, dt, threads, pm = cpu_fit_once(a, phis, a.cpu_rows, budget)
        tot_models += m
        tot_s += dt
        pmax = max(pmax, pm)
    v = tot_models / tot_s
    line = {'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': a.gpus, 'steps': a.steps,
            'warmup': a.warmup, 'ms_per_step': 1e3 * tot_s / a.steps, 'higher_is_better': True, 'scaling': 'strong',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': config_dict(a, n_total, 1, {'cpu_sample_rows': a.cpu_rows}),
            'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                             'sample': cpu_sample_desc(a, a.cpu_rows, budget, tot_models // max(a.steps, 1), pmax)},
            'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


def load_phis(a):
    import warnings
    from FoKL import getKernels
    if bench_data.CONFIGS[a.workload]['kernel'] == 'Cubic Splines':
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            return getKernels.sp500()
    return getKernels.bernoulli()


# ---------------------------------------------------------------------------------------------------
def measure_fp64_peak(torch, dev):
    """cuBLAS DGEMM ceiling on this box (MEASURED_PEAKS.json has no FP64 figure): 4096^3, best of 5."""
    n = 4096
    A = torch.randn(n, n, dtype=torch.float64, device=dev)
    B = torch.randn(n, n, dtype=torch.float64, device=dev)
    torch.matmul(A, B)
    best = float('inf')
    for _ in range(5):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        torch.matmul(A, B)
        e.record()
        torch.cuda.synchronize(dev)
        best = min(best, s.elapsed_time(e))
    del A, B
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def main():
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
        return
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (the B200 path has no CPU fallback); use --impl reference for the CPU arm')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    from FoKL import FoKLRoutines as FR

    c = bench_data.CONFIGS[a.workload]
    n_total = a.n or c['n']
    per = -(-n_total // world)
    lo, hi = rank * per, min((rank + 1) * per, n_total)
    x_host, y_host = bench_data.make_rows(a.workload, lo, hi, n_total=n_total)
    # pinned host buffers (the e2e arm copies from these every step)
    x_pin = torch.from_numpy(x_host).pin_memory()
    y_pin = torch.from_numpy(y_host).pin_memory()
    x_host, y_host = x_pin.numpy(), y_pin.numpy()
    phis = load_phis(a)
    m = c['m']
    unit_minmax = [[0.0, 1.0]] * m      # the generator draws inputs in [0, 1): global bounds, identical on every rank

    eng = FR._engine()
    eng.set_phis(phis, c['kernel'])

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def new_model():
        return bench_data.make_model(FR, a.workload, draws=a.draws, phis=phis)

    ds = eng.upload(x_host, y_host)

    def step_resident():
        np.random.seed(c['seed'])
        model = new_model()
        model.fit(ds, None)
        return dict(FR.LAST_FIT_INFO)

    def step_e2e():
        np.random.seed(c['seed'])
        model = new_model()
        betas, mtx, evs = model.fit(x_host, y_host, clean=True, minmax=unit_minmax, AutoTranspose=False)
        info = dict(FR.LAST_FIT_INFO)
        info['d2h'] = betas.nbytes + mtx.nbytes + evs.nbytes
        return info

    def timed(fn, steps, profile):
        barrier()
        eng.profile = {} if profile else None
        l0 = eng.launch_count()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        s.record()
        infos = [fn() for _ in range(steps)]
        e.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = s.elapsed_time(e)
        t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        prof = eng.profile_summary() if profile else {}
        eng.profile = None
        return infos, float(t[0].item()), float(t[1].item()), eng.launch_count() - l0, prof

    fp64_peak = measure_fp64_peak(torch, dev) if rank == 0 else None

    for _ in range(a.warmup):
        step_resident()
    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
    infos, ms, wall_ms, launches, prof = timed(step_resident, a.steps, True)
    clk = clocks.stop() if rank == 0 else None
    models = sum(i['n_gibbs'] for i in infos)
    value = models / (ms * 1e-3)

    e2e = None
    if not a.no_e2e:
        step_e2e()
        einfos, ems, ewall, _, _ = timed(step_e2e, a.steps, False)
        emodels = sum(i['n_gibbs'] for i in einfos)
        e2e = {'value': emodels / (ems * 1e-3), 'unit': UNIT,
               'h2d_bytes_per_step': int(x_host.nbytes + y_host.nbytes), 'd2h_bytes_per_step': int(einfos[-1]['d2h']),
               'ms_per_step': ems / a.steps}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    hbm_src = 'MEASURED_PEAKS.json hbm_gbs (of measured)' if 'hbm_gbs' in peaks else '6650 GB/s (of fallback)'

    def stage(name):
        return prof.get(name, {'ms': 0.0, 'calls': 0})

    # DRAM traffic per launch from the committed `ncu --set full` capture of this same command (tools/ncu_traffic.py
    # over the .ncu-rep of one cfg4 fit); a number taken under the profiler is evidence, not a timing
    traffic = {}
    try:
        with open(os.path.join(ROOT, 'profiles', 'r01_ncu_traffic.json')) as f:
            traffic = json.load(f)
    except Exception:
        pass

    def traffic_of(kernel):
        t = traffic.get(kernel)
        if not t or a.workload != 'cfg4' or a.n or world != 1:
            return None
        return t['traffic_bytes_per_launch']

    tot_ms = ms
    shares = {k: v['ms'] / tot_ms for k, v in prof.items()}
    b = stage('basis')
    roof = None
    if b['calls']:
        ach = b['bytes'] / (b['ms'] * 1e-3) / 1e9
        roof = {'kernel': 'basis_kernel (K1)', 'bound': 'hbm', 'achieved': ach, 'peak': hbm_peak, 'unit': 'GB/s',
                'frac': ach / hbm_peak, 'traffic': traffic_of('basis_kernel'), 'peak_source': hbm_src,
                'launches': b['calls'],
                'avg_launch_ms': b['ms'] / b['calls'], 'bytes_per_launch': b['bytes'] / b['calls'],
                'share_of_step': shares.get('basis')}
    g = stage('gram')
    roof_g = None
    if g['calls'] and fp64_peak:
        ach = g['flops'] / (g['ms'] * 1e-3) / 1e12
        roof_g = {'kernel': 'gram_kernel (K2, FP64 DMMA)', 'bound': 'tensor', 'achieved': ach, 'peak': fp64_peak,
                  'unit': 'TFLOP/s', 'frac': ach / fp64_peak, 'traffic': traffic_of('gram_kernel'),
                  'peak_source': 'in-run cuBLAS DGEMM 4096^3 (torch.matmul f64), best of 5', 'launches': g['calls'],
                  'avg_launch_ms': g['ms'] / g['calls'], 'share_of_step': shares.get('gram')}

    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        mdl, dt, threads, pm = cpu_fit_once(a, phis, a.cpu_rows, a.cpu_seconds)
        cpu = {'value': mdl / dt, 'unit': UNIT, 'cores': threads, 'kind': 'port',
               'sample': cpu_sample_desc(a, a.cpu_rows, a.cpu_seconds, mdl, pm), 'seconds': dt}

    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': a.steps, 'warmup': a.warmup,
            'ms_per_step': ms / a.steps, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic', 'config': config_dict(a, n_total, world),
            'candidate_models_per_step': models / a.steps, 'terms_selected': infos[-1]['terms'],
            'substages': infos[-1]['substages'], 'wall_ms_per_step': wall_ms / a.steps,
            'e2e': e2e, 'gpu_launches': int(launches), 'clocks': clk, 'roofline': roof, 'roofline_gram': roof_g,
            'roofline_note': "roofline = the basis-matrix kernel BASELINE.json's metric names (HBM-bound); "
                             "roofline_gram = the kernel with the largest share of the step (FP64 tensor pipe)",
            'stage_ms_per_step': {k: v['ms'] / a.steps for k, v in prof.items()},
            'cpu_baseline': cpu}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
