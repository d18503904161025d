"""Layer pipeline across the GPUs of one box: host-side logic.

The reference has no multi-GPU mode -- ggml offloads layers of ONE context to ONE device (rwkv.cpp:97-116). RWKV's natural
shard is a contiguous block of layers per GPU: rank g keeps layers [g*L/N, (g+1)*L/N) with their weights AND their slice of
the recurrent state resident, stage 0 also owns the embedding, the last stage ln_out + head. The only traffic is the
activation hand-off at each stage boundary: x f32[C x T] (+ v_first for v7), 16 KB per token at 7B.

A single stream cannot go faster this way (layer i+1 needs layer i), so throughput comes from keeping N sequences in flight,
one per stage (``schedule``): at tick t rank r works on sequence (t - r) mod N. In steady state every rank is busy and one
token (or one prefill chunk) leaves the last stage per tick.

Everything here is transport-agnostic and GPU-free so it is covered by world_size-2 gloo tests on CPU
(tests/test_pipeline_host.py); bench.py plugs in the CUDA stage (rwkv_b200_stage_eval on the NCCL stream) and torch.distributed
send/recv over NVLink.
"""
from typing import Callable, List, Optional, Sequence, Tuple


def stage_layers(n_layer: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block of layers of `rank`: as even as possible, earlier ranks take the remainder (their extra embedding
    gather is cheap, the last rank also runs the head)."""
    if not (0 <= rank < world) or world < 1 or n_layer < world:
        raise ValueError(f"cannot split {n_layer} layers over {world} stages (rank {rank})")
    base, extra = divmod(n_layer, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def stage_layers_balanced(n_layer: int, world: int, rank: int, head_layers: float = 0.0, embed_layers: float = 0.0) -> Tuple[int, int]:
    """Contiguous blocks that minimise the heaviest stage when the last stage also streams the head (`head_layers` = head bytes /
    bytes of one layer: 3.2 at RWKV-6-7B with an FP16 head) and the first one gathers embeddings (`embed_layers`, ~0). A pipeline
    ticks at the pace of its slowest stage, so an even split of the LAYERS leaves N - 1 stages idle while the last one reads the
    head. Exact min-max partition by dynamic programming over the cut positions; every stage keeps at least one layer."""
    if not (0 <= rank < world) or world < 1 or n_layer < world:
        raise ValueError(f"cannot split {n_layer} layers over {world} stages (rank {rank})")
    INF = float("inf")

    def cost(stage: int, a: int, b: int) -> float:
        return (b - a) + (head_layers if stage == world - 1 else 0.0) + (embed_layers if stage == 0 else 0.0)
    # best[s][e]: smallest possible maximum over stages 0..s when stage s ends at layer e
    best = [[INF] * (n_layer + 1) for _ in range(world)]
    cut = [[0] * (n_layer + 1) for _ in range(world)]
    for e in range(1, n_layer + 1):
        best[0][e] = cost(0, 0, e)
    for st in range(1, world):
        for e in range(st + 1, n_layer + 1):
            for a in range(st, e):
                v = max(best[st - 1][a], cost(st, a, e))
                if v < best[st][e] - 1e-12:
                    best[st][e], cut[st][e] = v, a
    bounds, e = [n_layer], n_layer
    for st in range(world - 1, 0, -1):
        e = cut[st][e]
        bounds.append(e)
    bounds.append(0)
    bounds.reverse()
    return bounds[rank], bounds[rank + 1]


def schedule(world: int, n_sequences: int, n_items: int, rank: int) -> List[Optional[Tuple[int, int]]]:
    """What `rank` does at each of the n_items + world - 1 ticks: None while the pipeline fills / drains, else (sequence, step).
    Work item u (u = 0 .. n_items - 1) is step u // n_sequences of sequence u % n_sequences; it enters stage 0 at tick u and reaches
    stage r at tick u + r, so a sequence never occupies two stages at once when n_sequences >= world."""
    if n_sequences < world:
        raise ValueError("need at least one in-flight sequence per stage")
    out: List[Optional[Tuple[int, int]]] = []
    for t in range(n_items + world - 1):
        u = t - rank
        out.append((u % n_sequences, u // n_sequences) if 0 <= u < n_items else None)
    return out


def hidden_floats(n_embed: int, n_tokens: int, arch_major: int) -> int:
    """Size of one hand-off: x, plus v_first for v7 (rwkv_b200_stage_hidden_len)."""
    return (2 if arch_major == 7 else 1) * n_embed * n_tokens


class Transport:
    """Point-to-point hand-off between neighbouring stages over torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""

    def __init__(self, dist, rank: int, world: int):
        self.dist, self.rank, self.world = dist, rank, world

    def recv_from_prev(self, buf) -> None:
        if self.rank > 0:
            self.dist.recv(buf, src=self.rank - 1)

    def send_to_next(self, buf) -> None:
        if self.rank < self.world - 1:
            self.dist.send(buf, dst=self.rank + 1)


def run_ticks(transport: Transport, plan: Sequence[Optional[Tuple[int, int]]], stage_fn: Callable, recv_buf, send_buf) -> int:
    """Drives one rank through `plan`: receive the activations of the work item (ranks > 0), run the stage, pass its output on
    (ranks < world - 1). stage_fn(sequence, step, hidden_in, hidden_out) runs this rank's layers; on GPUs it only enqueues work on
    the stream the transport uses, so receive -> compute -> send are ordered on the device without host synchronisation.
    Returns the number of work items processed."""
    done = 0
    first, last = transport.rank == 0, transport.rank == transport.world - 1
    for item in plan:
        if item is None:
            continue
        seq, step = item
        if not first:
            transport.recv_from_prev(recv_buf)
        stage_fn(seq, step, None if first else recv_buf, None if last else send_buf)
        if not last:
            transport.send_to_next(send_buf)
        done += 1
    return done
