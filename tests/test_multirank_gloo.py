"""World-size-2 gloo test (CPU) of the N>1 host logic: every replicate and every speculative search
step is owned by exactly one rank, and applying the stop rule in step order to gathered results gives
the sequential answer."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sequential_stop(d2, p_th, p_step=0.001, tank_n=5):
    """The reference's loop (:725-761) on a precomputed sequence of second-smallest values."""
    p_, tank = 0.999, []
    for step, d in enumerate(d2):
        tank.append(d)
        last = tank[-tank_n:]
        if (sum(x < p_th for x in last) > tank_n - 1) or p_ < 0.9:
            return step + 1, p_ + (tank_n - 1) * p_step
        p_ -= p_step
    raise AssertionError("sequence too short")


def worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import ctypes as C
    import torch
    from sclens_b200 import _lib
    from sclens_b200._lib import ptr
    from sclens_b200.api import speculative_search_merge
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = _lib.load()
    ids = np.empty(64, np.int32)
    cnt = C.c_int32()
    assert lib.scl_plan_replicates(21, world, rank, ptr(ids, C.c_int32), C.byref(cnt)) == 0
    mine = torch.zeros(21, dtype=torch.int32)
    mine[ids[:cnt.value].tolist()] = 1
    dist.all_reduce(mine)
    assert bool((mine == 1).all())                      # each replicate owned exactly once
    # speculative search: rank g evaluates step wave*world+g of a fixed synthetic d2 sequence
    rng = np.random.default_rng(0)
    d2_all = np.concatenate([0.3 - 0.01 * np.arange(12), 0.1 + 0.001 * rng.standard_normal(30)])
    p_th = 0.15
    gathered, stop = [], None
    wave = 0
    while stop is None:
        step = C.c_int32()
        assert lib.scl_plan_search_wave(wave, world, rank, C.byref(step)) == 0
        local = torch.tensor([d2_all[step.value]], dtype=torch.float64)
        out = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(out, local)
        gathered += [float(o.item()) for o in out]
        stop = speculative_search_merge(gathered, p_th)
        wave += 1
    want = sequential_stop(d2_all, p_th)
    assert stop[0] == want[0] and stop[1] == want[1], (stop, want)
    q.put((rank, stop))
    dist.destroy_process_group()


def test_two_rank_partition_and_speculative_search():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    res = sorted(q.get() for _ in range(2))
    assert res[0][1] == res[1][1]


def test_gram_shard_plan_partitions_the_contraction_axis():
    """scl_plan_gram_shard (the block of cells / genes a rank densifies and contracts before the ncclAllReduce): for
    any length and world size the blocks are 8-aligned, ordered, disjoint and cover the padded axis exactly once."""
    import ctypes as C
    sys.path.insert(0, ROOT)
    from sclens_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(0)
    lengths = list(range(1, 700, 7)) + [20000, 68000, 500000] + [int(x) for x in rng.integers(1, 2_000_000, 200)]
    for K in lengths:
        ld = (K + 7) // 8 * 8
        for world in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(world):
                k0, k1 = C.c_int64(), C.c_int64()
                assert lib.scl_plan_gram_shard(K, world, r, C.byref(k0), C.byref(k1)) == 0
                a, b = k0.value, k1.value
                assert a % 8 == 0 and b % 8 == 0 and 0 <= a <= b <= ld, (K, world, r, a, b)
                if b > a:
                    assert a == prev, (K, world, r, a, prev)      # contiguous with the previous non-empty block
                    prev = b
            assert prev == ld, (K, world, prev, ld)
    assert lib.scl_plan_gram_shard(0, 2, 0, C.byref(C.c_int64()), C.byref(C.c_int64())) != 0
    assert lib.scl_plan_gram_shard(10, 2, 2, C.byref(C.c_int64()), C.byref(C.c_int64())) != 0


def test_pass_task_plan_deals_every_solve_once_and_in_order():
    """scl_plan_pass_task (the wave schedule of scl_run_pass): tasks 0 (data), 1 (null), 2 (reference basis) and every search
    step are owned by exactly one (wave, rank); the reference basis never lands in a later wave than the first search step,
    and a host simulation of the waves - stop rule fed in task order - ends exactly like the sequential loop."""
    import ctypes as C
    sys.path.insert(0, ROOT)
    from sclens_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(1)
    d2_all = np.concatenate([0.3 - 0.01 * np.arange(9), 0.1 + 0.001 * rng.standard_normal(40)])
    p_th = 0.15
    want = sequential_stop(d2_all, p_th)
    for world in (1, 2, 3, 4, 8):
        seen, step_of, steps = {}, {}, []
        for wave in range(6):
            for r in range(world):
                t, s = C.c_int32(), C.c_int32()
                assert lib.scl_plan_pass_task(wave, world, r, C.byref(t), C.byref(s)) == 0
                assert t.value not in seen
                seen[t.value] = (wave, r)
                step_of[(wave, r)] = s.value
                if s.value >= 0:
                    steps.append(s.value)
        assert sorted(seen) == list(range(6 * world))
        assert steps == list(range(len(steps)))                   # every step once, in task order
        refine_slot = [t for t in range(3, 6 * world) if step_of[seen[t]] < 0]
        # with three ranks or more the refinement of the data spectrum takes the first slot of wave 1: rank 0's, which holds
        # the data matrix's eigenvectors and has touched nothing else since
        assert refine_slot == ([world] if world >= 3 else []) and all(seen[t] == (1, 0) for t in refine_slot)
        first_step_wave = min(w for (w, r), s in step_of.items() if s == 0)
        assert seen[2][0] <= first_step_wave and seen[0][0] <= seen[2][0] and seen[1][0] <= seen[2][0]
        # waves until the stop rule fires: steps of a wave are consumed in task order
        p_, tank, n_steps, stop, wave = 0.999, [], 0, None, 0
        while stop is None:
            for r in range(world):
                t, sv = C.c_int32(), C.c_int32()
                assert lib.scl_plan_pass_task(wave, world, r, C.byref(t), C.byref(sv)) == 0
                s = sv.value
                if s < 0 or stop is not None:
                    continue
                tank.append(d2_all[s])
                n_steps += 1
                if (sum(x < p_th for x in tank[-5:]) > 4) or p_ < 0.9:
                    stop = (n_steps, p_ + 4 * 0.001)
                    break
                p_ -= 0.001
            wave += 1
        assert stop[0] == want[0] and stop[1] == want[1], (world, stop, want)
    assert lib.scl_plan_pass_task(0, 2, 2, C.byref(C.c_int32()), None) != 0
