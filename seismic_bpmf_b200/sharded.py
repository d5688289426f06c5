"""Multi-GPU sharding of the two hot paths (one process per GPU, torch.distributed plumbing).

* matched filter: templates are independent -> block-partition T over the ranks, replicate the
  traces, one ``all_gather`` of the per-shard (T/G, n_corr) CC slabs at the end (SURVEY.md 8e;
  station chunks are additive, BPMF/similarity_search.py:517-538, but summing them would need an
  all-reduce of the whole slab, 8x the traffic).
* beamforming, reduce="max": sources are independent up to the final max -> block-partition K,
  every rank reports global source ids, one ``all_gather`` of (max, argmax) pairs followed by a
  local G-way select with ties to the lowest source index.

The compute step is injectable (``compute=``) so the partition/gather/merge logic is exercised on
CPU with the ``gloo`` backend in tests; the default compute is the CUDA path (no CPU fallback).
"""
import torch
import torch.distributed as dist


def shard_bounds(n, rank, world):
    """Contiguous block partition of range(n): first (n % world) ranks get one extra item."""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _world(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def _mf_compute_cuda(templates, moveouts, weights, data, step, normalize, precision):
    from . import device as dev

    return dev.matched_filter_dev(templates, moveouts, weights, data, step=step, normalize=normalize,
                                  network_sum=True, precision=precision)


def _matched_filter_sharded_overlapped(templates, moveouts, weights, data, step, normalize, precision, group,
                                       batch):
    """CUDA path with the exchange hidden behind the compute: the local shard runs in template batches and
    every finished batch is pushed to the peers' buffers by the copy engines (PeerSlabGather) while the next
    batch computes; the batches are then laid out in template order."""
    from . import device as dev

    rank, world = _world(group)
    T, S, C, L = templates.shape
    lo, hi = shard_bounds(T, rank, world)
    rows = (T + world - 1) // world  # longest shard: every rank runs the same number of batches
    plan = dev.MatchedFilterPlan(max(hi - lo, 1), S, C, L, data.shape[-1], step=step, normalize=normalize,
                                 precision=precision, device=data.device, batch=batch)
    n_corr, pitch = plan.n_corr, plan.pitch
    gather = PeerSlabGather(batch * pitch, data.device, group)
    out = torch.empty((T, n_corr), dtype=torch.float32, device=data.device)
    tp, mv, w = templates[lo:hi].contiguous(), moveouts[lo:hi].contiguous(), weights[lo:hi].contiguous()
    if hi > lo:
        plan.prepare(data)
    empty = torch.zeros(0, dtype=torch.float32, device=data.device)
    for b, t0 in enumerate(range(0, rows, batch)):
        n_here = max(0, min(batch, (hi - lo) - t0))
        if n_here > 0:
            plan.run_batch(tp, mv, w, data, t0, n_here)
            gather.push(b % 2, plan.cc[t0: t0 + n_here].reshape(-1))
        else:
            gather.push(b % 2, empty)  # short shard: still take part in the batch barrier
        # Lay the batch out in template order RIGHT BEHIND its own barrier on the ordering stream.  A peer's push
        # of batch b + 2 into this same buffer waits for barrier b + 1 of its own ordering stream, and that
        # barrier completes only once this rank has arrived at it -- i.e. after this read: no peer, however far
        # ahead, can overwrite the buffer while it is still being read here.
        _scatter_gathered(out, gather, b % 2, t0, T, world, pitch, n_corr)
    torch.cuda.current_stream().wait_stream(gather.stream)
    return out


def _scatter_gathered(out, gather, buf, t0, T, world, pitch, n_corr):
    """Copy batch rows [t0, t0 + batch) of every rank's shard from gather buffer ``buf`` into template order
    (on the gather's ordering stream, behind that batch's barrier)."""
    with torch.cuda.stream(gather.stream):
        view = gather.local[buf].view(world, -1)
        for r in range(world):
            a, b_ = shard_bounds(T, r, world)
            n = max(0, min(b_ - a - t0, view.shape[1] // pitch))
            if n > 0:
                out[a + t0: a + t0 + n] = view[r, : n * pitch].view(n, pitch)[:, :n_corr]


def matched_filter_sharded(templates, moveouts, weights, data, step=1, normalize="short",
                           precision="auto", group=None, gather=True, compute=None, overlap=True, batch=64):
    """Template-sharded matched filter.

    Every rank passes the FULL (T,S,C,L) templates / (T,S,C) moveouts, weights and the (S,C,N)
    traces (torch tensors on its own device).  Returns the (T, n_corr) CC sums on every rank when
    ``gather`` is true, else ``(local_cc, (t_lo, t_hi))``.  On CUDA with more than one rank the exchange is
    overlapped with the compute (``overlap``); otherwise one NCCL / gloo all-gather follows the compute.
    """
    rank, world = _world(group)
    T = templates.shape[0]
    lo, hi = shard_bounds(T, rank, world)
    if gather and overlap and compute is None and world > 1 and data.is_cuda:
        try:
            return _matched_filter_sharded_overlapped(templates, moveouts, weights, data, step, normalize,
                                                      precision, dist.group.WORLD if group is None else group, batch)
        except (ImportError, AttributeError, RuntimeError) as exc:  # symmetric memory unavailable
            import warnings

            warnings.warn("overlapped exchange unavailable (%s); falling back to an NCCL all-gather" % (exc,))
    compute = compute or _mf_compute_cuda
    n_corr = (data.shape[-1] - templates.shape[-1]) // step + 1
    if hi > lo:
        local = compute(templates[lo:hi].contiguous(), moveouts[lo:hi].contiguous(),
                        weights[lo:hi].contiguous(), data, step, normalize, precision)
    else:
        local = torch.zeros((0, n_corr), dtype=torch.float32, device=data.device)
    if not gather:
        return local, (lo, hi)
    if world == 1:
        return local
    # equal-sized slabs for the collective: pad the short shards with one zero row
    rows = (T + world - 1) // world
    slab = torch.zeros((rows, n_corr), dtype=torch.float32, device=data.device)
    slab[: hi - lo] = local
    full = torch.empty(world * rows * n_corr, dtype=torch.float32, device=data.device)
    dist.all_gather_into_tensor(full, slab.view(-1), group=group)
    full = full.view(world * rows, n_corr)
    out = torch.empty((T, n_corr), dtype=torch.float32, device=data.device)
    for r in range(world):
        a, b = shard_bounds(T, r, world)
        out[a:b] = full[r * rows: r * rows + (b - a)]
    return out


def merge_max_argmax(maxes, args):
    """(G,N) per-shard maxima and GLOBAL source ids -> (N,), (N,) with ties to the lowest id."""
    best = maxes.max(dim=0).values
    big = torch.iinfo(args.dtype).max
    cand = torch.where(maxes == best.unsqueeze(0), args, torch.full_like(args, big))
    return best, cand.min(dim=0).values


def _beam_compute_cuda(features, moveouts, weights_phases, weights_sources, out_of_bounds, source_offset):
    from . import device as dev

    K, S, P = moveouts.shape
    plan = dev.BeamformPlan(K, S, features.shape[1], P, features.shape[-1], device=features.device)
    return plan.run_max(features, moveouts, weights_phases, weights_sources, out_of_bounds, source_offset)


def beamform_max_sharded(features, moveouts, weights_phases, weights_sources, out_of_bounds="strict",
                         group=None, compute=None):
    """Source-sharded beamforming with reduce="max"; returns (maxbeam, maxbeam_sources) on every rank."""
    rank, world = _world(group)
    K = moveouts.shape[0]
    N = features.shape[-1]
    lo, hi = shard_bounds(K, rank, world)
    compute = compute or _beam_compute_cuda
    if hi > lo:
        mb, arg = compute(features, moveouts[lo:hi].contiguous(), weights_phases,
                          weights_sources[lo:hi].contiguous(), out_of_bounds, lo)
    else:  # an empty shard must never win: -inf, id = int32 max
        mb = torch.full((N,), float("-inf"), dtype=torch.float32, device=features.device)
        arg = torch.full((N,), torch.iinfo(torch.int32).max, dtype=torch.int32, device=features.device)
    if world == 1:
        return mb, arg
    packed_m = torch.empty(world * N, dtype=torch.float32, device=features.device)
    packed_a = torch.empty(world * N, dtype=torch.int32, device=features.device)
    dist.all_gather_into_tensor(packed_m, mb.contiguous().view(-1), group=group)
    dist.all_gather_into_tensor(packed_a, arg.contiguous().view(-1), group=group)
    packed_m, packed_a = packed_m.view(world, N), packed_a.view(world, N)
    if features.is_cuda:
        from . import device as dev

        best, barg = packed_m[0].clone(), packed_a[0].clone()
        for r in range(1, world):
            dev.beam_merge_(best, barg, packed_m[r], packed_a[r])
        return best, barg
    return merge_max_argmax(packed_m, packed_a)


class PeerSlabGather:
    """All-gather of equally sized float32 slabs through NVLink/NVSwitch peer memory, moved by the copy
    engines instead of SM kernels.

    The matched-filter kernel is persistent and owns every SM (one CTA each, all of its shared memory), so
    an NCCL kernel enqueued beside it only runs in its tail.  Here every rank PUSHES its finished slab
    straight into the gather buffers of its peers (torch symmetric memory gives the peer mappings;
    ``cudaMemcpyAsync`` on a side stream runs on the DMA engines) while the next template batch computes; a
    device-side barrier per batch tells every rank that the batch is complete everywhere.
    """

    def __init__(self, n_elements_per_rank, device, group=None, n_buffers=2, n_streams=3):
        import torch.distributed._symmetric_memory as symm_mem

        group = dist.group.WORLD if group is None else group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.n = int(n_elements_per_rank)
        self.stream = torch.cuda.Stream(device=device)  # ordering stream: barriers + the wait() handle
        # copies to different peers go out on several streams so that more than one copy engine is busy
        self.copy_streams = [torch.cuda.Stream(device=device) for _ in range(max(1, min(n_streams, self.world - 1)))]
        self._copy_events = [torch.cuda.Event() for _ in self.copy_streams]
        self.local, self.peers, self.handles = [], [], []
        for _ in range(n_buffers):
            buf = symm_mem.empty(self.world * self.n, dtype=torch.float32, device=device)
            handle = symm_mem.rendezvous(buf, group.group_name)
            self.local.append(buf)
            self.handles.append(handle)
            self.peers.append([handle.get_buffer(r, (self.world * self.n,), torch.float32)
                               for r in range(self.world)])
        self._event = torch.cuda.Event()

    def push(self, b, slab_flat):
        """Copy this rank's slab (flat, <= n elements, produced on the current stream) into slot ``rank`` of
        buffer ``b`` on every rank; returns immediately (work runs on the side stream)."""
        n = slab_flat.numel()
        assert n <= self.n and slab_flat.is_contiguous()
        self._event.record()
        lo = self.rank * self.n
        for i, cs in enumerate(self.copy_streams):
            cs.wait_event(self._event)
            cs.wait_stream(self.stream)  # keep batches ordered behind the previous barrier
            with torch.cuda.stream(cs):
                for k in range(i, self.world, len(self.copy_streams)):
                    r = (self.rank + k) % self.world  # own copy first, then the ring of peers
                    self.peers[b][r][lo: lo + n].copy_(slab_flat, non_blocking=True)
                self._copy_events[i].record()
        with torch.cuda.stream(self.stream):
            for ev in self._copy_events:
                self.stream.wait_event(ev)
            self.handles[b].barrier(channel=0)  # every rank's push of this batch has landed everywhere

    def gathered(self, b, rows, pitch):
        """View of buffer ``b`` as (world, rows, pitch) once the side stream has been waited for."""
        return self.local[b][: self.world * self.n].view(self.world, -1)[:, : rows * pitch].view(self.world, rows, pitch)

    def wait(self):
        """Make the current stream wait for every push/barrier enqueued so far."""
        torch.cuda.current_stream().wait_stream(self.stream)
