"""Sharding of the hot path over the GPUs of one box (one process per GPU, torch.distributed / NCCL over NVLink).

Two levels of independence exist:
  * text lines are fully independent (reference test_sr.py:77): shard lines, no collective;
  * characters are independent inside TSPGAN (reference networks.py:134-164 has no cross-sample op): shard the
    characters of a batch over ranks and all-gather the two prior tensors the SR decoder needs for its per-character
    concat (fea64 [n,256,64,64] + fea32 [n,512,32,32] = 6 MiB fp32 per character) -- the only exchange on the path.
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous, balanced [begin, end) of n items for `rank` (first n % world ranks get one extra)."""
    base, extra = divmod(n, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_lines(num_lines, rank=None, world=None):
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    return range(*shard_range(num_lines, rank, world))


def _all_gather_rows(local, counts, group):
    """All-gather tensors whose first dimension differs per rank (padded to the largest shard, then trimmed)."""
    world = len(counts)
    mx = max(counts)
    pad = local
    if local.shape[0] < mx:
        pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[:local.shape[0]] = local
    pad = pad.contiguous()
    out = torch.empty((world * mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    if all(c == mx for c in counts):
        return out
    return torch.cat([out[r * mx:r * mx + counts[r]] for r in range(world)], dim=0)


def generate_priors_sharded(generator, styles, labels, group=None, pipeline_chunks=1):
    """Character-sharded TSPGAN: every rank generates its contiguous shard of the characters and the priors are
    all-gathered so that every rank holds the full (image, fea64, fea32) -- what TSPSRNet consumes.

    `generator(styles, labels, None) -> (image, fea64, fea32)` is the TSPGAN module (or any callable with its contract);
    tensors are returned in the generator's own memory format (channels_last views stay channels_last).

    pipeline_chunks > 1 (equal shards only): the local shard is generated in that many sub-chunks and the NCCL all-gather of
    sub-chunk i runs asynchronously while sub-chunk i+1 is generated (6 MiB per character, ~0.75 TB/s over NVLink).
    Measured on 8 x B200 (profiles/r1_priors_sharded_8gpu*.json): no gain (29.0 vs 27.8 ms for 1024 characters) -- the conv
    kernels are persistent with one CTA per SM, so NCCL's copy kernels get no SM until a conv kernel retires.  Hiding the
    exchange needs copy-engine / peer-store transfers (symmetric memory), which is next-round work; default stays 1."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = labels.shape[0]
    if world == 1:
        return generator(styles, labels, None)
    if pipeline_chunks > 1 and n % (world * pipeline_chunks) == 0:
        return _generate_pipelined(generator, styles, labels, group, world, rank, pipeline_chunks)
    counts = [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]
    b, e = shard_range(n, rank, world)
    if e > b:
        outs = generator(styles[b:e], labels[b:e], None)
    else:   # more ranks than characters: contribute an empty shard of the right trailing shape
        probe = generator(styles[:1], labels[:1], None)
        outs = tuple(o[:0] for o in probe)
    gathered = []
    for o in outs:
        cl = o.dim() == 4 and o.permute(0, 2, 3, 1).is_contiguous()          # NHWC storage behind an NCHW-shaped view
        local = o.permute(0, 2, 3, 1) if cl else o.contiguous()
        full = _all_gather_rows(local, counts, group)
        gathered.append(full.permute(0, 3, 1, 2) if cl else full)
    return tuple(gathered)


def _generate_pipelined(generator, styles, labels, group, world, rank, chunks):
    """Block-cyclic sharding: sub-chunk c covers the characters [c*world*sub, (c+1)*world*sub) and rank r generates the
    r-th block of `sub` of them, so every sub-chunk's all-gather output is one contiguous slice of the final tensors (no
    re-ordering copy) and its NCCL transfer can overlap the generation of sub-chunk c+1."""
    n = labels.shape[0]
    sub = n // (world * chunks)
    full, works = None, []
    for c in range(chunks):
        b = c * world * sub + rank * sub
        outs = generator(styles[b:b + sub], labels[b:b + sub], None)
        if full is None:
            full = [torch.empty((n,) + tuple(o.permute(0, 2, 3, 1).shape[1:]), dtype=o.dtype, device=o.device) for o in outs]
        for o, buf in zip(outs, full):
            local = o.permute(0, 2, 3, 1).contiguous()
            dst = buf[c * world * sub:(c + 1) * world * sub]
            works.append((dist.all_gather_into_tensor(dst, local, group=group, async_op=True), local))
    for w, _ in works:
        w.wait()
    return tuple(buf.permute(0, 3, 1, 2) for buf in full)


# ---------------------------------------------------------------------------------------------------------------------
# Owner-only exchange (round 2).  The all-gather above gives EVERY rank ALL priors (6 MiB per character: 6.6 GB received per rank
# for 1024 characters at 8 GPUs, measured 8.9 ms at the NVLink all-gather limit, after the compute) although only the rank that
# runs the SR decoder of a line ever reads that line's priors (reference consumer: networks.py:442-445, 475-478).  Here lines are
# owned by ranks (contiguous blocks), characters are generated BLOCK-CYCLICALLY -- rank r generates the r-th sub-block of every
# owner's characters -- and one all-to-all delivers each sub-block to its owner, where it lands in natural character order:
# (world-1)/world of 6 MiB per OWNED character crosses NVLink instead of (world-1) x 6 MiB.
# ---------------------------------------------------------------------------------------------------------------------
def owner_blocks(n_chars, world):
    """Block-cyclic plan for n_chars = world*world*sub characters.  Returns sub; owner o owns [o*world*sub, (o+1)*world*sub),
    and rank r generates, for every owner o, characters [o*world*sub + r*sub, +sub)."""
    if n_chars % (world * world) != 0:
        raise ValueError(f"owner exchange needs the character count ({n_chars}) to be a multiple of world^2 = {world * world}")
    return n_chars // (world * world)


def generate_priors_for_owners(generator, styles, labels, group=None, keep=(1, 2), exchange=True):
    """Character-sharded TSPGAN with owner-only exchange.  ``styles`` / ``labels`` describe ALL characters (identical on every
    rank); returns the priors of the characters this rank OWNS (``[rank*n/world, (rank+1)*n/world)``, natural order) as a tuple
    of the generator outputs selected by ``keep`` (default: fea64, fea32 -- what the SR decoder reads; the 128-px image stays
    where it was generated), each a channels_last NCHW-shaped view like the generator's own outputs.

    exchange=False skips the all-to-all and returns the LOCALLY GENERATED characters instead (bench: compute-only time).
    One NCCL all_to_all_single per kept tensor, equal splits of ``sub`` characters."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = labels.shape[0]
    if world == 1:
        outs = generator(styles, labels, None)
        return tuple(outs[k] for k in keep)
    sub = owner_blocks(n, world)
    idx = torch.cat([torch.arange(o * world * sub + rank * sub, o * world * sub + (rank + 1) * sub) for o in range(world)])
    idx_s = idx.to(styles.device)
    outs = generator(styles.index_select(0, idx_s), labels.index_select(0, idx.to(labels.device)), None)
    result = []
    for k in keep:
        o = outs[k]
        cl = o.dim() == 4 and o.permute(0, 2, 3, 1).is_contiguous()
        local = o.permute(0, 2, 3, 1) if cl else o.contiguous()            # [world*sub, ...]: block d goes to owner d
        if not exchange:
            result.append(o)
            continue
        recv = torch.empty_like(local)                                      # block s arrives from rank s = sub-block s of my range
        dist.all_to_all_single(recv, local.contiguous(), group=group)
        result.append(recv.permute(0, 3, 1, 2) if cl else recv)
    return tuple(result)


def exchange_bytes_per_rank(n_chars, world, bytes_per_char=6 * (1 << 20)):
    """Bytes each rank sends (= receives) in generate_priors_for_owners vs in the all-gather variant."""
    owned = n_chars // world
    return dict(all_to_all=owned * bytes_per_char * (world - 1) // world, all_gather=(n_chars - owned) * bytes_per_char)


# ---------------------------------------------------------------------------------------------------------------------
# In-kernel exchange (round 2): the convolution that produces a feature tap stores it -- from its own epilogue, tile by tile,
# while the MMAs of the next tile run -- straight into the symmetric-memory receive buffer of the rank that owns the character's
# line (mn_conv_params.y2_ptrs; NVLink peer stores).  No separate collective moves prior features any more; one device-side
# barrier per round publishes them.  Same block-cyclic plan as generate_priors_for_owners.
# ---------------------------------------------------------------------------------------------------------------------
class PeerPriorExchange:
    """Receive buffers in symmetric memory (torch.distributed._symmetric_memory: cuMem allocations mapped into every rank of
    the box) + the per-character destination pointer tables of one block-cyclic round of ``n_chars`` characters.

    >>> ex = PeerPriorExchange(n_chars=1024, device=dev)            # collective: every rank constructs it
    >>> f64, f32 = ex.generate(tspgan, styles_all, labels_all)      # this rank's OWNED characters, natural order

    ``slots`` receive buffers are used round-robin so that the SR decoder may still read round k while round k+1 is written."""

    F64 = 64 * 64 * 256
    F32 = 32 * 32 * 512

    def __init__(self, n_chars, device, group=None, slots=2):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.n, self.sub = n_chars, owner_blocks(n_chars, self.world)
        self.own = n_chars // self.world
        self.slots, self.slot = slots, 0
        self.device = torch.device(device)
        per_slot = self.own * (self.F64 + self.F32)
        self.buf = symm_mem.empty(slots * per_slot, dtype=torch.float32, device=self.device)
        self.handle = symm_mem.rendezvous(self.buf, self.group)
        peers = [int(p) for p in self.handle.buffer_ptrs]
        w, r, sub = self.world, self.rank, self.sub
        self.ptrs = []
        for s in range(slots):
            t64, t32 = [], []
            for d in range(w):                                   # local sample j = d*sub + i goes to owner d, position r*sub + i
                base64 = peers[d] + 4 * s * per_slot
                base32 = base64 + 4 * self.own * self.F64
                for i in range(sub):
                    t64.append(base64 + 4 * (r * sub + i) * self.F64)
                    t32.append(base32 + 4 * (r * sub + i) * self.F32)
            self.ptrs.append({64: torch.tensor(t64, dtype=torch.int64).to(self.device), 32: torch.tensor(t32, dtype=torch.int64).to(self.device)})
        self.idx = torch.cat([torch.arange(o * w * sub + r * sub, o * w * sub + (r + 1) * sub) for o in range(w)])
        self.idx_dev = self.idx.to(self.device)
        self.per_slot = per_slot

    def local_views(self, slot):
        """(fea64 [own,256,64,64], fea32 [own,512,32,32]) NCHW-shaped channels_last views of this rank's receive buffer."""
        base = slot * self.per_slot
        f64 = self.buf[base:base + self.own * self.F64].view(self.own, 64, 64, 256).permute(0, 3, 1, 2)
        f32 = self.buf[base + self.own * self.F64:base + self.per_slot].view(self.own, 32, 32, 512).permute(0, 3, 1, 2)
        return f64, f32

    def generate(self, generator, styles, labels):
        """``styles`` / ``labels``: ALL n characters (identical on every rank).  Every rank generates its block-cyclic share; the tap
        convolutions deliver the features to the owners; returns the OWNED characters' (fea64, fea32), valid until ``slots``
        further rounds have been generated."""
        slot = self.slot
        self.slot = (self.slot + 1) % self.slots
        lab = labels.index_select(0, self.idx_dev if labels.is_cuda else self.idx)
        generator(styles.index_select(0, self.idx_dev), lab, None, _tap_ptrs=self.ptrs[slot])
        self.handle.barrier(channel=slot)                        # device-side, stream-ordered: every rank's stores have landed
        return self.local_views(slot)
