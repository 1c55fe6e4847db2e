"""Multi-GPU layout of the path: one process per GPU, weights replicated, whole utterances sharded
across ranks, NO collective on the inference data path - the only exchange is the host-side gather of the finished
per-utterance results (and, in training, the gradient all-reduce in `train.py`).

The reference has no multi-GPU inference at all (its drivers walk the utterance list in one process,
convert.py:224-265, 342-360); what has to be preserved is what a single process would have produced:
  * results in utterance order, and
  * the Gumbel draw of every chunk (model/model.py:96: `torch.rand` on the CPU generator, in call order) - so every
    rank draws the noise of ALL chunks in the global order and keeps the draws of its own utterances.
"""
import numpy as np
import torch
import torch.distributed as dist

from .frontend import MIN_LEN, one_hot_rows, segment_plan
from .model import Encoder


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


def output_lengths(n_frames, seg_len):
    """(unit frames, spectrogram frames) convert()/encode() produce for an utterance of `n_frames` frames - known on every
    rank from the length alone (convert.py:139-168: one model call per chunk, the tail chunk drops a frame; the decoder
    emits 8 frames per unit frame; padded utterances keep MIN_LEN // 8 units)."""
    _, plan, keep = segment_plan(int(n_frames), seg_len)
    t8 = [Encoder.t8(e - s) for s, e in plan]
    return (sum(t8) if keep is None else keep), 8 * sum(t8)


def gather_arrays(local, shards, shapes, dtype, group=None, device=None):
    """Tensor-collective gather of per-utterance arrays: `local[k]` is this rank's array for utterance `shards[rank][k]`,
    `shapes[i]` the shape every rank can compute for utterance i.  Each rank ships ONE flat buffer (padded to the
    largest rank's size) through `all_gather_into_tensor` - no pickling, no per-object messages - and cuts the others'
    buffers by the known shapes.  Returns the arrays of all utterances in order (numpy)."""
    on = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    rank = dist.get_rank(group) if on else 0
    world = len(shards)
    sizes = [[int(np.prod(shapes[i])) for i in part] for part in shards]
    totals = [sum(p) for p in sizes]
    flat = np.concatenate([np.asarray(a, dtype).reshape(-1) for a in local]) if local else np.zeros(0, dtype)
    if flat.size != totals[rank]:
        raise RuntimeError(f'rank {rank}: {flat.size} elements produced, {totals[rank]} expected from the utterance lengths')
    if on:
        pad = max(totals)
        dev = device if device is not None else ('cuda' if dist.get_backend(group) == 'nccl' else 'cpu')
        mine = torch.zeros(pad, dtype=torch.from_numpy(flat[:0]).dtype, device=dev)
        mine[:flat.size] = torch.from_numpy(flat).to(dev)
        allbuf = torch.empty(world * pad, dtype=mine.dtype, device=dev)
        dist.all_gather_into_tensor(allbuf, mine, group=group)
        allnp = allbuf.cpu().numpy().reshape(world, pad)
    else:
        allnp = flat.reshape(1, -1)
    out = [None] * len(shapes)
    for r, part in enumerate(shards):
        off = 0
        for i, n in zip(part, sizes[r]):
            out[i] = allnp[r, off:off + n].reshape(shapes[i])
            off += n
    if any(o is None for o in out):
        raise RuntimeError('an utterance was produced by no rank')
    return out


class ShardedPath:
    """`AutoencoderPath` over the ranks of a process group: same calls, same results as one process.

    Units travel as int32 ids (4 bytes per unit frame) and results are exchanged with tensor collectives of known
    shapes (`gather_arrays`); `gather=False` keeps each rank's results local (per-rank file output, the layout of
    `test_encode`, convert.py:342-360, where every utterance becomes its own file anyway).
    Noise: `reference_noise_order=True` reproduces one process bit for bit - every rank then draws the noise of ALL chunks
    from torch's CPU generator (an O(total chunks) serial section per rank: ~60 us per 128-frame chunk); `noise_seed=<int>`
    draws on the device, one stream per (utterance, chunk): no serial section, and still independent of the world size."""

    def __init__(self, path, group=None):
        self.path, self.group = path, group
        on = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if on else 0
        self.world = dist.get_world_size(group) if on else 1

    def shards(self, specs):
        return shard_utterances([len(s) for s in specs], self.world, self.path.seg_len)

    def _units(self, units, specs, shards, as_ids):
        enc = self.path.Encoder
        lens = [output_lengths(len(s), self.path.seg_len)[0] for s in specs]
        if enc.enc_mode == 'one_hot':
            ids = gather_arrays(units, shards, [(n,) for n in lens], np.int32, self.group)
            return ids if as_ids else [one_hot_rows(i, enc.enc_size) for i in ids]
        return gather_arrays(units, shards, [(n, enc.enc_size) for n in lens], np.float32, self.group)

    def encode_utterances(self, specs, reference_noise_order=True, as_ids=False, noise_seed=None, gather=True):
        shards = self.shards(specs)
        mine = shards[self.rank]
        units = self.path.encode_utterances(specs, reference_noise_order, only=mine,
                                            as_ids=self.path.Encoder.enc_mode == 'one_hot', noise_seed=noise_seed)
        if not gather:
            return mine, units
        return self._units(units, specs, shards, as_ids)

    def convert_utterances(self, specs, target_speakers, enc_only=True, reference_noise_order=True, as_ids=False,
                           noise_seed=None, gather=True):
        shards = self.shards(specs)
        mine = shards[self.rank]
        out, units = self.path.convert_utterances(specs, target_speakers, enc_only, reference_noise_order, only=mine,
                                                  as_ids=self.path.Encoder.enc_mode == 'one_hot', noise_seed=noise_seed)
        if not gather:
            return mine, out, units
        c_out = self.path.Decoder.c_out
        spec_shapes = [(output_lengths(len(s), self.path.seg_len)[1], c_out) for s in specs]
        return gather_arrays(out, shards, spec_shapes, np.float32, self.group), self._units(units, specs, shards, as_ids)
