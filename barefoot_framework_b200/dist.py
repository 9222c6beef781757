"""Multi-GPU plumbing for the hot path: one process per GPU (torchrun), no data-path collective.

The reference shards work items evenly over SLURM nodes and collects one small record per item
through pickle files (barefoot.py:2200-2209, subProcess.py:96-168).  Here the hyper-parameter sets
(or, when there are fewer sets than ranks, the candidate tiles) are split across ranks and only
the per-set (value, index) records are exchanged with torch.distributed (NCCL on GPUs; the same
code runs on gloo/CPU tensors in the tests)."""
import torch
import torch.distributed as dist

_BIG = 2 ** 62


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def broadcast_object(obj, src=0):
    """Rank `src`'s Python object on every rank (host RNG state, wall-clock costs, stop decisions of the
    orchestrator): the reference has ONE driver process, so everything that steers the control flow must
    come from one rank.  Identity in a single process."""
    r, w = world()
    if w == 1:
        return obj
    box = [obj if r == src else None]
    if dist.get_backend() == "nccl":
        dist.broadcast_object_list(box, src=src, device=torch.device("cuda", torch.cuda.current_device()))
    else:
        dist.broadcast_object_list(box, src=src)
    return box[0]


def raise_together(err, what="stage"):
    """Combine a per-rank failure (None or the exception a rank caught while computing its shard) BEFORE the
    stage's collective: if any rank failed, every rank raises - the failing ranks their own exception, the
    others a RuntimeError naming the first failing rank - so no rank is left waiting in an all-gather for a
    peer that has already unwound (george raises LinAlgError from one work item; the reference's pool then
    fails the whole map, barefoot.py:1070-1073).  One all-reduce of a single integer."""
    r, w = world()
    if w == 1:
        if err is not None:
            raise err
        return
    flag = torch.tensor([w if err is None else r], dtype=torch.int64)
    if dist.get_backend() == "nccl":
        flag = flag.to(torch.device("cuda", torch.cuda.current_device()))
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    first = int(flag.item())
    if err is not None:
        raise err
    if first < w:
        raise RuntimeError("%s failed on rank %d (see that rank's exception); raised on every rank so that "
                           "no collective is left half-entered" % (what, first))


def gather_host_sets(values, indices, n_total=None):
    """gather_sets for host arrays (numpy in, numpy out); identity in a single process."""
    r, w = world()
    if w == 1:
        return values, indices
    import numpy as np
    tv = torch.as_tensor(np.ascontiguousarray(values, dtype=np.float64))
    ti = torch.as_tensor(np.ascontiguousarray(indices, dtype=np.int64))
    if dist.get_backend() == "nccl":            # NCCL exchanges device tensors only
        dev = torch.device("cuda", torch.cuda.current_device())
        tv, ti = tv.to(dev), ti.to(dev)
    v, i = gather_sets(tv, ti, n_total)
    return v.cpu().numpy(), i.cpu().numpy()


def shard_range(n, rank=None, world_size=None):
    """Contiguous, balanced split of n units: returns (start, stop) of this rank
    (the first n % world ranks get one extra unit)."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    base, extra = divmod(int(n), world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def _comm(t):
    """gloo (CPU tests, or ranks sharing one GPU) exchanges host tensors; NCCL device tensors."""
    backend = dist.get_backend()
    if backend == "gloo" and t.is_cuda:
        return t.cpu()
    if backend == "nccl" and not t.is_cuda:
        return t.to(torch.device("cuda", torch.cuda.current_device()))
    return t


def gather_sets(values, indices, n_total=None):
    """Set-sharded stages: every rank holds (values [h_r], indices [h_r]) for its contiguous block of
    hyper-parameter sets; returns the concatenation over ranks in set order on every rank.  With
    n_total (the number of sets before sharding with shard_range) every rank knows all block sizes and
    the exchange is ONE all-gather of a padded [2, m] block (indices travel bit-cast to float64);
    without it the sizes are exchanged first."""
    r, w = world()
    if w == 1:
        return values, indices
    dev = values.device
    if n_total is not None:
        sizes = [b - a for a, b in (shard_range(n_total, q, w) for q in range(w))]
        if values.numel() != sizes[r] or indices.numel() != sizes[r]:
            raise ValueError("gather_sets: rank %d holds %d sets, shard_range(%d) gives %d"
                             % (r, values.numel(), n_total, sizes[r]))
        m = max(sizes)
        pack = torch.zeros((2, m), dtype=torch.float64, device=dev)
        pack[0, : values.numel()] = values
        pack[1, : indices.numel()] = indices.to(torch.int64).view(torch.float64)
        pack = _comm(pack).view(-1)
        out = torch.empty((w * 2 * m,), dtype=torch.float64, device=pack.device)
        dist.all_gather_into_tensor(out, pack)
        out = out.view(w, 2, m)
        gv = torch.cat([out[q, 0, :s] for q, s in enumerate(sizes)])
        gi = torch.cat([out[q, 1, :s] for q, s in enumerate(sizes)]).view(torch.int64)
        return gv.to(dev), gi.to(dev)
    values, indices = _comm(values), _comm(indices)
    n_local = torch.tensor([values.numel()], dtype=torch.int64, device=values.device)
    sizes = [torch.zeros_like(n_local) for _ in range(w)]
    dist.all_gather(sizes, n_local)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes)
    pv = torch.zeros((m,), dtype=values.dtype, device=values.device)
    pi = torch.zeros((m,), dtype=indices.dtype, device=indices.device)
    pv[: values.numel()] = values
    pi[: indices.numel()] = indices
    gv = [torch.empty_like(pv) for _ in range(w)]
    gi = [torch.empty_like(pi) for _ in range(w)]
    dist.all_gather(gv, pv)
    dist.all_gather(gi, pi)
    return (torch.cat([g[:s] for g, s in zip(gv, sizes)]).to(dev), torch.cat([g[:s] for g, s in zip(gi, sizes)]).to(dev))


def _gather_packed(*cols):
    """ONE all-gather of several equally long float64 / int64 vectors (int64 travels bit-cast to float64):
    returns one [world, len] tensor per input, on the communication device."""
    r, w = world()
    n = cols[0].numel()
    pack = torch.empty((len(cols), n), dtype=torch.float64, device=cols[0].device)
    for k, c in enumerate(cols):
        pack[k] = c.view(torch.float64) if c.dtype == torch.int64 else c
    pack = _comm(pack).view(-1)
    out = torch.empty((w * len(cols) * n,), dtype=torch.float64, device=pack.device)
    dist.all_gather_into_tensor(out, pack)
    out = out.view(w, len(cols), n)
    return [out[:, k, :].contiguous().view(torch.int64) if c.dtype == torch.int64 else out[:, k, :]
            for k, c in enumerate(cols)]


class _ShardReduction:
    """Handle of a candidate-shard reduction whose all-gather is still in flight (NCCL runs it on its own stream
    after the producing kernels): result() makes the current stream wait for it and applies the first-index
    maximum with one small kernel.  Lets the next step's set-up and sweep start behind the collective."""

    def __init__(self, work, pack, gathered, world_size, S, device):
        # `pack` is read by the collective on NCCL's stream: it must stay allocated until the collective has run
        # (a freed block is handed to the NEXT work item's tensors on the compute stream, whose kernels are not
        # ordered against NCCL's read - found by tools/dbg/pipeline_stress.py: records of neighbouring items mixed)
        self.work, self.pack, self.gathered = work, pack, gathered
        self.world_size, self.S, self.device = world_size, S, device

    def result(self):
        from . import _lib
        if self.work is not None:
            self.work.wait()
        best_val = torch.empty((self.S,), dtype=torch.float64, device=self.device)
        best_idx = torch.empty((self.S,), dtype=torch.int64, device=self.device)
        _lib.call("bf_records_reduce", self.world_size, self.S, self.gathered.data_ptr(), best_val.data_ptr(),
                  best_idx.data_ptr(), _lib.stream())
        # the caller's stream now orders everything after the collective; the buffers may be released once the
        # reduce kernel has run - record that stream so the allocator waits for it
        self.gathered.record_stream(torch.cuda.current_stream())
        self.pack = None
        return best_val, best_idx


class _Ready:
    def __init__(self, out):
        self.out = out

    def result(self):
        return self.out


def reduce_candidate_shards(values, indices, index_offset, async_op=False):
    """Candidate-sharded stages (fewer sets than ranks): every rank holds the per-set maximum over
    ITS candidates, (values [S], local indices [S]).  Returns the maximum over all candidates with
    numpy's first-index tie rule (np.where(v == max)[0][0] over the concatenated candidate array,
    acquisitionFunc.py:135-138): among equal values the lowest GLOBAL index wins.  One collective; on CUDA
    tensors over NCCL the records are packed and reduced by one small kernel each (bf_records_pack /
    bf_records_reduce).  async_op=True returns a handle whose result() gives the pair, so the caller can launch
    more work behind the collective."""
    r, w = world()
    if w == 1:
        out = (values, indices + int(index_offset))
        return _Ready(out) if async_op else out
    dev = values.device
    if values.is_cuda and dist.get_backend() == "nccl":
        from . import _lib
        S = values.numel()
        pack = torch.empty((2 * S,), dtype=torch.float64, device=dev)
        _lib.call("bf_records_pack", S, values.contiguous().data_ptr(), indices.to(torch.int64).contiguous().data_ptr(),
                  int(index_offset), pack.data_ptr(), _lib.stream())
        gathered = torch.empty((w * 2 * S,), dtype=torch.float64, device=dev)
        work = dist.all_gather_into_tensor(gathered, pack, async_op=True)
        handle = _ShardReduction(work, pack, gathered, w, S, dev)
        return handle if async_op else handle.result()
    gidx = indices + int(index_offset)
    gv, gi = _gather_packed(values.contiguous(), gidx.to(torch.int64).contiguous())      # [world, S] each
    best = gv.max(dim=0).values
    cand = torch.where(gv == best, gi, torch.full_like(gi, _BIG))
    out = (best.to(dev), cand.min(dim=0).values.to(dev))
    return _Ready(out) if async_op else out


def reduce_top2(t1, i1, t2, index_offset):
    """Candidate-sharded KG: combine per-rank (largest mean, its first global index, second largest)
    into the global triple that bf_kg_diag needs.  One collective."""
    r, w = world()
    gi1 = i1 + int(index_offset)
    if w == 1:
        return t1, gi1, t2
    dev = t1.device
    a, b, c = _gather_packed(t1.contiguous(), gi1.to(torch.int64).contiguous(), t2.contiguous())   # [world, S]
    best = a.max(dim=0).values
    is_best = a == best
    first = torch.where(is_best, b, torch.full_like(b, _BIG)).min(dim=0).values
    # second largest "by position": every first-place value except the winning one, and all seconds
    winner = is_best & (b == first)
    others = torch.where(winner, torch.full_like(a, float("-inf")), a)
    second = torch.maximum(others.max(dim=0).values, c.max(dim=0).values)
    return best.to(dev), first.to(dev), second.to(dev)


def gather_rows(local, n_total=None):
    """Row-sharded vectors (k-medoids row sums): concatenation of the ranks' contiguous shard_range
    blocks.  With n_total given every rank knows all block sizes, so this is ONE all-gather of equally
    padded blocks and no host synchronisation."""
    r, w = world()
    if w == 1:
        return local
    if n_total is None:
        out, _ = gather_sets(local, torch.zeros((local.numel(),), dtype=torch.int64, device=local.device))
        return out
    dev = local.device
    sizes = [shard_range(n_total, q, w) for q in range(w)]
    m = max(b - a for a, b in sizes)
    pad = torch.zeros((m,), dtype=local.dtype, device=local.device)
    pad[: local.numel()] = local
    pad = _comm(pad)
    out = torch.empty((w * m,), dtype=local.dtype, device=pad.device)
    dist.all_gather_into_tensor(out, pad)
    out = out.view(w, m)
    return torch.cat([out[q, : b - a] for q, (a, b) in enumerate(sizes)]).to(dev)


def all_gather_flat(out, local):
    """out [world, ...] <- every rank's `local` (equal shapes), one collective; gloo exchanges host copies
    (CPU tests, ranks sharing one GPU)."""
    r, w = world()
    if w == 1:
        out.copy_(local.reshape(out.shape))
        return out
    if dist.get_backend() == "gloo" and local.is_cuda:
        h = torch.empty(out.shape, dtype=out.dtype)
        dist.all_gather_into_tensor(h.view(-1), local.cpu().contiguous().view(-1))
        out.copy_(h)
        return out
    dist.all_gather_into_tensor(out.view(-1), local.contiguous().view(-1))
    return out


def sum_parts(t):
    """In-place sum over ranks of a vector whose every element is non-zero on exactly one rank (the
    in-cluster cost parts of k-medoids); adding zeros is exact, so the result is rank-count invariant."""
    r, w = world()
    if w == 1:
        return t
    if dist.get_backend() == "gloo" and t.is_cuda:
        h = t.cpu()
        dist.all_reduce(h, op=dist.ReduceOp.SUM)
        t.copy_(h)
        return t
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def gather_row_blocks(local, rows_per_rank):
    """Concatenate 2-D float64 blocks [r_q, C] held by the ranks (r_q = rows_per_rank[q], known to every
    rank from shard_range) in rank order with ONE all-gather of equally padded blocks.  Used by the ROM
    stage, whose (model, fantasy point) groups are split over the ranks."""
    r, w = world()
    if w == 1:
        return local
    if local.dim() != 2 or local.shape[0] != rows_per_rank[r]:
        raise ValueError("gather_row_blocks: rank %d holds a block of shape %s, expected %d rows"
                         % (r, tuple(local.shape), rows_per_rank[r]))
    dev, C = local.device, local.shape[1]
    m = max(rows_per_rank)
    if m == 0:
        return local
    pad = torch.zeros((m, C), dtype=torch.float64, device=dev)
    pad[: local.shape[0]] = local
    pad = _comm(pad).view(-1)
    out = torch.empty((w * m * C,), dtype=torch.float64, device=pad.device)
    dist.all_gather_into_tensor(out, pad)
    out = out.view(w, m, C)
    return torch.cat([out[q, : rows_per_rank[q]] for q in range(w)]).to(dev)
