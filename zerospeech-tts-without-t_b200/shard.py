"""Multi-GPU layout of the path: one process per GPU, weights replicated, whole utterances sharded
across ranks, NO collective on the inference data path - the only exchange is the host-side gather of the finished
per-utterance results (and, in training, the gradient all-reduce in `train.py`).

The reference has no multi-GPU inference at all (its drivers walk the utterance list in one process,
convert.py:224-265, 342-360); what has to be preserved is what a single process would have produced:
  * results in utterance order, and
  * the Gumbel draw of every chunk (model/model.py:96: `torch.rand` on the CPU generator, in call order) - so every
    rank draws the noise of ALL chunks in the global order and keeps the draws of its own utterances.
"""
import torch
import torch.distributed as dist

from .frontend import segment_plan


def segments_of(n_frames, seg_len):
    """Number of model calls the reference makes for an utterance of `n_frames` frames (convert.py:139-165)."""
    return len(segment_plan(n_frames, seg_len)[1])


def shard_utterances(lengths, world, seg_len=128):
    """Whole utterances -> ranks, balanced on frames the model sees (longest-first greedy, deterministic).
    Returns `world` sorted index lists that partition range(len(lengths))."""
    if world < 1:
        raise ValueError('world must be >= 1')
    cost = []
    for i, n in enumerate(lengths):
        padded, plan, _ = segment_plan(int(n), seg_len)
        cost.append((sum(e - s for s, e in plan), i))
    load = [0] * world
    shards = [[] for _ in range(world)]
    for c, i in sorted(cost, key=lambda t: (-t[0], t[1])):
        r = min(range(world), key=lambda k: (load[k], k))
        load[r] += c
        shards[r].append(i)
    return [sorted(s) for s in shards]


def gather_in_order(local, mine, n_total, group=None):
    """`local[k]` is this rank's result for utterance `mine[k]`; returns the list of all `n_total` results in
    utterance order on every rank (host objects through `all_gather_object`; single process: a re-ordering)."""
    if len(local) != len(mine):
        raise ValueError('one result per owned utterance expected')
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        pairs = [list(zip(mine, local))]
    else:
        pairs = [None] * dist.get_world_size(group)
        dist.all_gather_object(pairs, list(zip(mine, local)), group=group)
    out = [None] * n_total
    seen = 0
    for part in pairs:
        for i, r in part:
            if out[i] is not None:
                raise RuntimeError(f'utterance {i} was produced by two ranks')
            out[i] = r
            seen += 1
    if seen != n_total:
        raise RuntimeError(f'{n_total - seen} utterances were produced by no rank')
    return out


def reference_order_noise(lengths, seg_len, noise_shape, sampler, keep):
    """Draws the Gumbel noise of every chunk of every utterance in the reference's call order (one `sampler(shape)`
    call per chunk, convert.py:203-213 -> model/model.py:96) and returns {utterance: [noise per chunk]} for the
    utterances in `keep`; the other draws only advance the generator."""
    keep = set(keep)
    out = {}
    for u, n in enumerate(lengths):
        _, plan, _ = segment_plan(int(n), seg_len)
        draws = [sampler(noise_shape(1, e - s)) for s, e in plan]
        if u in keep:
            out[u] = draws
    return out


class ShardedPath:
    """`AutoencoderPath` over the ranks of a process group: same calls, same results as one process."""

    def __init__(self, path, group=None):
        self.path, self.group = path, group
        on = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if on else 0
        self.world = dist.get_world_size(group) if on else 1

    def _mine(self, specs):
        return shard_utterances([len(s) for s in specs], self.world, self.path.seg_len)[self.rank]

    def encode_utterances(self, specs, reference_noise_order=True):
        mine = self._mine(specs)
        units = self.path.encode_utterances(specs, reference_noise_order, only=mine)
        return gather_in_order(units, mine, len(specs), self.group)

    def convert_utterances(self, specs, target_speakers, enc_only=True, reference_noise_order=True):
        mine = self._mine(specs)
        out, units = self.path.convert_utterances(specs, target_speakers, enc_only, reference_noise_order, only=mine)
        both = gather_in_order(list(zip(out, units)), mine, len(specs), self.group)
        return [b[0] for b in both], [b[1] for b in both]
