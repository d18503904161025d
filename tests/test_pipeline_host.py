"""Host logic of the layer pipeline (rwkv.cpp_b200/pipeline.py) on CPU: layer partition, tick schedule, and a
world_size-2 gloo run in which each rank applies its block of "layers" of a toy recurrent model and hands the activations on --
the result must equal the single-process evaluation of all layers, sequence by sequence (state stays on its stage)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__  # noqa: E402

pipeline = __graft_entry__.load_package().pipeline


def test_stage_layers_cover_every_layer_once():
    for n_layer in (12, 24, 32, 33):
        for world in (1, 2, 4, 8):
            if n_layer < world:
                continue
            blocks = [pipeline.stage_layers(n_layer, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n_layer
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in blocks]
            assert max(sizes) - min(sizes) <= 1
    assert pipeline.stage_layers(32, 8, 3) == (12, 16)
    with pytest.raises(ValueError):
        pipeline.stage_layers(4, 8, 0)


def test_balanced_partition_accounts_for_the_head():
    """The last stage also streams the head (3.2 layers' worth of bytes at RWKV-6-7B): blocks are contiguous, cover every layer once,
    and the heaviest stage is as light as a contiguous partition allows (checked against brute force on small cases)."""
    import itertools
    for n_layer, world, head in ((32, 8, 3.2), (32, 4, 3.2), (32, 2, 3.2), (24, 8, 1.5), (12, 4, 6.0), (12, 4, 0.0), (9, 3, 2.0)):
        blocks = [pipeline.stage_layers_balanced(n_layer, world, r, head_layers=head) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n_layer and all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
        assert all(e > b for b, e in blocks)
        got = max((e - b) + (head if r == world - 1 else 0.0) for r, (b, e) in enumerate(blocks))
        if n_layer <= 12:
            best = min(max((c[i + 1] - c[i]) + (head if i == world - 1 else 0.0) for i in range(world))
                       for cuts in itertools.combinations(range(1, n_layer), world - 1) for c in [(0,) + cuts + (n_layer,)])
            assert abs(got - best) < 1e-9, (n_layer, world, head, got, best)
    assert [pipeline.stage_layers_balanced(32, 8, r, head_layers=3.2) for r in range(8)][-1] == (31, 32)
    assert pipeline.stage_layers_balanced(12, 4, 2, head_layers=0.0) == (6, 9)


def test_schedule_keeps_a_sequence_on_one_stage_at_a_time():
    world, n_seq, items = 4, 4, 16
    plans = [pipeline.schedule(world, n_seq, items, r) for r in range(world)]
    ticks = items + world - 1
    assert all(len(p) == ticks and sum(x is not None for x in p) == items for p in plans)
    for t in range(ticks):
        busy = [p[t][0] for p in plans if p[t] is not None]
        assert len(busy) == len(set(busy))                     # no sequence on two stages in the same tick
    # every (sequence, step) that enters stage 0 reaches the last stage `world - 1` ticks later
    for t in range(items):
        assert plans[-1][t + world - 1] == plans[0][t]
    assert plans[2][:2] == [None, None] and plans[0][-3:] == [None, None, None]      # pipeline fill / drain
    with pytest.raises(ValueError):
        pipeline.schedule(4, 2, 8, 0)


def test_hidden_floats():
    assert pipeline.hidden_floats(4096, 1, 6) == 4096
    assert pipeline.hidden_floats(2560, 128, 7) == 2 * 2560 * 128


N_LAYER, C, N_SEQ, STEPS = 6, 16, 2, 5


def toy_layer(i, x, state):
    """A stand-in for one RWKV layer: mixes the input with the layer's recurrent state and updates it."""
    w = np.cos(np.arange(C, dtype=np.float32) * (i + 1))
    y = np.tanh(x * w + 0.5 * state[i])
    state[i] = 0.9 * state[i] + 0.1 * y
    return x + y


def toy_input(seq, step):
    return np.sin(np.arange(C, dtype=np.float32) + 3.0 * seq + step).astype(np.float32)


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    begin, end = pipeline.stage_layers(N_LAYER, world, rank)
    states = [np.zeros((N_LAYER, C), np.float32) for _ in range(N_SEQ)]     # only rows [begin, end) are ever touched here
    results = {}
    recv, send = torch.zeros(C), torch.zeros(C)

    def stage(seq, step, hidden_in, hidden_out):
        x = toy_input(seq, step) if hidden_in is None else hidden_in.numpy().copy()
        for i in range(begin, end):
            x = toy_layer(i, x, states[seq])
        if hidden_out is None:
            results[(seq, step)] = x
        else:
            hidden_out.copy_(torch.from_numpy(x))

    plan = pipeline.schedule(world, N_SEQ, N_SEQ * STEPS, rank)
    n = pipeline.run_ticks(pipeline.Transport(dist, rank, world), plan, stage, recv, send)
    assert n == N_SEQ * STEPS
    if rank == world - 1:
        np.savez(os.path.join(out_dir, "out.npz"), **{f"{s}_{k}": v for (s, k), v in results.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_two_stage_pipeline_matches_single_process(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(os.path.join(str(tmp_path), "out.npz"))
    for seq in range(N_SEQ):
        state = np.zeros((N_LAYER, C), np.float32)
        for step in range(STEPS):
            x = toy_input(seq, step)
            for i in range(N_LAYER):
                x = toy_layer(i, x, state)
            np.testing.assert_array_equal(got[f"{seq}_{step}"], x)
