"""Multi-GPU checks of the library-issued collectives (run under torchrun on a box with >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py

Every rank must see (1) the row-sliced matvec + all-gather equal to the full single-GPU matvec, (2) the sliced-basis Lanczos
(tn_lanczos_lm1 with a tn_comm) return the eigenpair of the unsharded solver, bit-identical on all ranks, (3) the term-sharded
variant (all-reduce of H|psi>) agree as well, (4) a whole sharded DMRG run reproduce the reference golden, with ranks that
do NOT share a random seed (the initial state is broadcast)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    from tests.test_gpu_dmrg import para_from_golden
    from tnalg_b200 import Parameters as Pm, ops
    from tnalg_b200.DMRG_anyH import dmrg_finite_size
    from tnalg_b200.MPSClass import MpsOpenBoundaryClass
    be = ops.backend()
    comm = be.comm()
    assert comm is not None and comm.world == world and comm.rank == rank
    # ---- collectives ----
    t = torch.full((1000,), float(rank + 1), dtype=torch.float64, device=dev)
    comm.allreduce(t)
    assert float(t[0]) == world * (world + 1) / 2
    # small all-reduces (<= 64 doubles) go through the peer window when CUDA IPC works between the ranks: many in a row
    # (the slot ring is reused), interleaved with large NCCL ones
    base = torch.arange(22, dtype=torch.float64, device=dev)
    for it in range(40):
        s = base * (rank + 1) + it
        comm.allreduce(s)
        assert torch.equal(s, base * (world * (world + 1) / 2) + it * world), it
        if it % 13 == 0:
            big = torch.full((5000,), float(rank), dtype=torch.float64, device=dev)
            comm.allreduce(big)
            assert float(big[17]) == world * (world - 1) / 2
    if rank == 0:
        print('peer-window collectives: %d of %d' % (comm.peer_collectives(), comm.collectives()), flush=True)
    if not os.environ.get('TNALG_NO_PEER') and os.environ.get('TNALG_EXPECT_PEER'):
        assert comm.peer_collectives() >= 40
    bufs = [torch.full((64, 64), float(rank * 10 + j), dtype=torch.float64, device=dev) for j in range(5)]
    comm.broadcast_many(bufs, [j % world for j in range(5)])
    for j in range(5):
        assert float(bufs[j][3, 3]) == (j % world) * 10 + j
    be.shard_min_rows, be.shard_min_n = 4, 256      # slice even small sites in this check
    # ---- local problem: 4x3 lattice, chi = 96 ----
    para = Pm.generate_parameters_dmrg('square')
    para.update(square_width=4, square_height=3, chi=96, op=para['op'][:6], eigs_tol=1e-10)
    para = Pm.make_consistent_parameter_dmrg(para)
    np.random.seed(100 + rank)                       # different seeds on purpose
    A = MpsOpenBoundaryClass(para['l'], para['d'], para['chi'], operators=para['op'], is_save_op=True, eig_way=1)
    A.sync_replicas()
    p = 5
    A.shard_terms = False                            # gauge moves + environments replicated for this check
    A.correct_orthogonal_center(p)
    A.shard_terms = True
    args = (para['index1'], para['index2'], para['coeff1'], para['coeff2'])
    a, d, b = A.mps[p].shape
    full = A.effective_hamiltonian_plan(p, *args, tol=1e-10)
    x = A.mps[p].clone()
    want = full.matvec(x).clone()
    lam0, vec0, mv0, res0, ok0 = be.lanczos(full, para['tau'], x, 1e-10)
    rows = be.shard_rows(a, d, b, rank, world)
    assert rows is not None, (a, d, b)
    pr = A.effective_hamiltonian_plan(p, *args, tol=1e-10, rows=rows)
    got = pr.matvec(x)
    assert float((got - want[rows[0]:rows[0] + rows[1]]).abs().max()) <= 1e-13 * float(want.abs().max())
    lam1, vec1, mv1, res1, ok1 = be.lanczos(pr, para['tau'], x, 1e-10, comm=comm)
    assert ok0 and ok1 and abs(lam1 - lam0) < 1e-12, (lam0, lam1)
    assert abs(abs(float((vec0 * vec1).sum())) - 1) < 1e-9
    gathered = [torch.empty_like(vec1) for _ in range(world)]
    dist.all_gather(gathered, vec1)
    for g in gathered:
        assert bool((g == vec1).all())              # bit-identical on every rank
    pt = A.effective_hamiltonian_plan(p, *args, tol=1e-10, rank=rank, world=world)
    lam2, vec2, mv2, res2, ok2 = be.lanczos(pt, para['tau'], x, 1e-10, comm=comm)
    assert ok2 and abs(lam2 - lam0) < 1e-12 and abs(abs(float((vec0 * vec2).sum())) - 1) < 1e-9
    for pl in (full, pr, pt):
        pl.destroy()
    # ---- whole sharded runs against the reference goldens, both decompositions ----
    from tnalg_b200 import envs
    envs.EnvCache.shard_min_dim = 8                    # sharded bond moves (in-place all-gather of the new blocks) too
    for mode in ('rows', 'terms'):
        be.shard_mode = mode
        be.shard_min_rows, be.shard_min_n = 4, 256     # slice even the small sites of the golden cases
        for case in ('e2e_j1j2_4x2', 'e2e_chain12'):
            g = dict(np.load(os.path.join(ROOT, 'tests', 'golden', case + '.npz'), allow_pickle=False))
            para = para_from_golden(g)
            np.random.seed(int(g['seed']) if rank == 0 else 12345 + rank)
            ob, A, info, para = dmrg_finite_size(para)
            assert abs(ob['e_per_site'][0] - g['e_per_site'][0]) <= 1e-10 * abs(g['e_per_site'][0]), (mode, case)
            for k in ('eb_full', 'mx', 'mz', 'corr_z'):
                assert np.abs(np.asarray(ob[k]).reshape(-1) - g[k].reshape(-1)).max() < 1e-8, (mode, case, k)
            for n in range(para['l'] - 1):
                ref = g['lm_%d' % n]
                assert np.abs(A.lm[n] - ref).max() <= 1e-10 * ref.max() + 1e-12
            e_all = [None] * world
            dist.all_gather_object(e_all, float(ob['e_per_site'][0]))
            assert all(e == e_all[0] for e in e_all)
    # ---- odd slice lengths (spin-1, chi = 9: 5 rows of 27 numbers per rank): the ranks' window offsets differ in 16-byte
    # alignment, the arrival counts must not ----
    be.shard_mode = 'rows'
    be.shard_min_rows, be.shard_min_n = 4, 128
    e_odd = []
    for shard in (True, False):
        para = Pm.generate_parameters_dmrg('chain')
        para.update(spin='one', l=6, chi=9, eigs_tol=1e-12, shard_terms=shard)
        para = Pm.make_consistent_parameter_dmrg(para)
        np.random.seed(7)
        ob, A, info, para = dmrg_finite_size(para)
        e_odd.append(float(ob['e_per_site'][0]))
    assert abs(e_odd[0] - e_odd[1]) <= 1e-10 * abs(e_odd[1]), e_odd
    if rank == 0:
        print('multigpu_check ok: world %d, collectives issued by the library: %d (peer window: %d)' %
              (world, comm.collectives(), comm.peer_collectives()))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
