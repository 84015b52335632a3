"""Multi-GPU check (run on the GPU box): a row-sharded fit over WORLD_SIZE ranks must select the same terms and
reproduce the BIC trace of the single-GPU fit of the same dataset.

    python tools/dist_check.py --single                      # writes gpurun_out/dist_single.npz
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/dist_check.py                                  # compares against it on rank 0
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'fokl-gpy_b200'))
sys.path.insert(0, ROOT)
import bench_data  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--single', action='store_true')
    ap.add_argument('--cfg', default='cfg4')
    ap.add_argument('--n', '--rows', dest='n', type=int, default=1_000_000)   # use --rows under torchrun (its own parser claims --n*)
    ap.add_argument('--draws', type=int, default=500)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    world = 1 if a.single else int(os.environ.get('WORLD_SIZE', '1'))
    rank = 0 if a.single else int(os.environ.get('RANK', '0'))
    local = 0 if a.single else int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from FoKL import FoKLRoutines as FR
    c = bench_data.CONFIGS[a.cfg]
    per = -(-a.n // world)
    lo, hi = rank * per, min((rank + 1) * per, a.n)
    x, y = bench_data.make_rows(a.cfg, lo, hi, n_total=a.n)
    raw = 2.0 * x + 1.0                       # un-normalised inputs: exercises the MIN / MAX allreduce of clean()
    np.random.seed(c['seed'])
    model = bench_data.make_model(FR, a.cfg, draws=a.draws)
    betas, mtx, evs = model.fit(raw, y, clean=True)
    info = dict(FR.LAST_FIT_INFO)
    out = os.path.join(ROOT, 'gpurun_out')
    os.makedirs(out, exist_ok=True)
    if a.single:
        np.savez(os.path.join(out, 'dist_single.npz'), betas=betas, mtx=mtx, evs=evs, minmax=np.array(model.minmax))
        print('single: terms', mtx.shape, 'substages', len(evs), info)
        return
    # every rank must hold the same result
    t = torch.tensor([float(mtx.shape[0]), float(len(evs)), float(evs.sum()), float(np.abs(betas).sum())],
                     dtype=torch.float64, device='cuda')
    lo_t, hi_t = t.clone(), t.clone()
    dist.all_reduce(lo_t, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi_t, op=dist.ReduceOp.MAX)
    same = bool(torch.equal(lo_t, hi_t))
    if rank == 0:
        ref = np.load(os.path.join(out, 'dist_single.npz'))
        ok_terms = np.array_equal(ref['mtx'], mtx)
        ok_evs = len(ref['evs']) == len(evs) and np.allclose(ref['evs'], evs, rtol=1e-9, atol=0)
        ok_mm = np.allclose(ref['minmax'], np.array(model.minmax), rtol=0, atol=0)
        mean_ref, mean_now = ref['betas'].mean(axis=0), betas.mean(axis=0)
        ok_beta = mean_ref.shape == mean_now.shape and np.allclose(mean_ref, mean_now, rtol=1e-6, atol=1e-9 * np.abs(mean_ref).max())
        print('dist_check world=%d: ranks_identical=%s terms_equal=%s evs_equal=%s minmax_equal=%s posterior_means_equal=%s'
              % (world, same, ok_terms, ok_evs, ok_mm, ok_beta), info, flush=True)
        if not (same and ok_terms and ok_evs and ok_mm and ok_beta):
            print('evs single', ref['evs'], '\nevs dist  ', evs)
            sys.exit(1)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
