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


def gather_detections(series, index, value, thr, group=None, device=None):
    """All-gather of per-rank detection lists (numpy arrays of equal length per rank, GLOBAL series ids):
    one collective for the counts, one for the padded records.  Returns the concatenation sorted by
    (series, index) -- a few bytes per detection instead of the (T, n_corr) CC slab."""
    import numpy as np

    rank, world = _world(group)
    if world == 1:
        order = np.lexsort((index, series))
        return series[order], index[order], value[order], thr[order]
    device = torch.device("cpu") if device is None else device
    n = torch.tensor([len(series)], dtype=torch.int64, device=device)
    counts = torch.zeros(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(counts, n, group=group)
    counts = counts.cpu().numpy()
    m = int(counts.max())
    if m == 0:
        z = np.zeros(0)
        return z.astype(np.int64), z.astype(np.int32), z.astype(np.float32), z.astype(np.float32)
    # one (m, 4) float64 record per rank: series and index are exact in float64
    rec = np.zeros((m, 4), dtype=np.float64)
    rec[: len(series), 0], rec[: len(series), 1] = series, index
    rec[: len(series), 2], rec[: len(series), 3] = value, thr
    mine = torch.from_numpy(rec).to(device)
    everyone = torch.empty((world, m, 4), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(everyone.view(-1), mine.view(-1), group=group)
    everyone = everyone.cpu().numpy()
    rows = np.concatenate([everyone[r, : counts[r]] for r in range(world)], axis=0)
    order = np.lexsort((rows[:, 1], rows[:, 0]))
    rows = rows[order]
    return (rows[:, 0].astype(np.int64), rows[:, 1].astype(np.int32), rows[:, 2].astype(np.float32),
            rows[:, 3].astype(np.float32))


def matched_filter_detect_sharded(templates, moveouts, weights, data, gaussian_sample, num_dev, window_samp,
                                  overlap=0.25, search_win=0, step=1, normalize="short", precision="auto", group=None,
                                  batch=64, capacity=1 << 18, compute=None, detect=None, max_cc_threshold=None,
                                  max_kurtosis=None):
    """The multi-GPU detection path: every rank runs the matched filter on its block of templates, thresholds and
    declusters ITS OWN CC rows on the device (rms threshold of BPMF/libc.c:516-673 with ``half_window =
    window_samp // 2``, ``shift = int((1 - overlap) * window_samp)`` as BPMF/clib.py:296-297, then
    ``select_cc_indexes`` with ``search_win``; optionally the threshold cap ``max_cc_threshold * sum(weights)`` and
    the kurtosis sanity check of similarity_search.py:629-642), and only the detections are all-gathered.  The CC slab never
    leaves the GPU that made it and is never held for more than ``batch`` templates at a time, so
    BASELINE.json configs[3] (172.8 GB of CC) needs no more memory than one batch.

    Returns numpy ``(template, index, cc, threshold)`` sorted by (template, index), identical on every rank.
    ``compute`` / ``detect`` are injectable for the CPU (gloo) tests: ``compute`` as in matched_filter_sharded,
    ``detect(local_cc) -> (row, index, cc, threshold)``."""
    import numpy as np

    rank, world = _world(group)
    T, S, C, L = templates.shape
    lo, hi = shard_bounds(T, rank, world)
    half, shift = int(window_samp) // 2, int((1.0 - overlap) * int(window_samp))
    if compute is not None or detect is not None:
        assert compute is not None and detect is not None
        if hi > lo:
            local = compute(templates[lo:hi].contiguous(), moveouts[lo:hi].contiguous(), weights[lo:hi].contiguous(),
                            data, step, normalize, precision)
            s_, i_, v_, t_ = detect(local)
        else:
            s_, i_, v_, t_ = (np.zeros(0, np.int64), np.zeros(0, np.int32), np.zeros(0, np.float32),
                              np.zeros(0, np.float32))
        return gather_detections(np.asarray(s_, np.int64) + lo, i_, v_, t_, group, data.device)
    from . import device as dev

    n_local = hi - lo
    if n_local > 0:
        batch = max(1, min(int(batch), n_local))
        plan = dev.MatchedFilterPlan(batch, S, C, L, data.shape[-1], step=step, normalize=normalize,
                                     precision=precision, device=data.device, batch=batch)
        n_batches = (n_local + batch - 1) // batch
        det = dev.DetectBatches(n_batches, batch, plan.n_corr, half, shift, capacity=capacity, device=data.device)
        tp, mv, w = templates[lo:hi].contiguous(), moveouts[lo:hi].contiguous(), weights[lo:hi].contiguous()
        plan.prepare(data)
        for k, t0 in enumerate(range(0, n_local, batch)):
            n_here = min(batch, n_local - t0)
            plan.run_batch(tp, mv, w, data, t0, n_here, cc_row0=0)
            cap = None
            if max_cc_threshold is not None:
                cap = (float(max_cc_threshold) * w[t0: t0 + n_here].sum(dim=(1, 2))).float().contiguous()
            det.run(k, plan.cc[:n_here], gaussian_sample, num_dev, search_win=search_win, mode="select", thr_cap=cap,
                    max_kurtosis=max_kurtosis)
        s_, i_, v_, t_ = det.collect(series_offsets=[lo + k * batch for k in range(n_batches)])
    else:
        s_, i_, v_, t_ = (np.zeros(0, np.int64), np.zeros(0, np.int32), np.zeros(0, np.float32),
                          np.zeros(0, np.float32))
    return gather_detections(s_, i_, v_, t_, group, data.device)


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


def _beam_full_compute_cuda(features, moveouts, weights_phases, weights_sources, out_of_bounds):
    from . import device as dev

    K, S, P = moveouts.shape
    plan = dev.BeamformPlan(K, S, features.shape[1], P, features.shape[-1], device=features.device)
    return plan.run_full(features, moveouts, weights_phases, weights_sources, out_of_bounds)


def beamform_none_sharded(features, moveouts, weights_phases, weights_sources, out_of_bounds="strict", group=None,
                          gather=True, compute=None):
    """Source-sharded beamforming with ``reduce="none"`` (BPMF/template_search.py:560-569, consumed by
    BPMF/dataset.py:2184-2196): rank r stacks the block ``[lo, hi)`` of the K sources; the (K/G, N) blocks are
    concatenated in source order by one all-gather.  Returns the ``(K, N)`` beam on every rank
    when ``gather`` is true, else ``(local_beam, (lo, hi))`` -- the K x N array of a large grid only fits sharded."""
    rank, world = _world(group)
    K, N = moveouts.shape[0], features.shape[-1]
    lo, hi = shard_bounds(K, rank, world)
    compute = compute or _beam_full_compute_cuda
    if hi > lo:
        local = compute(features, moveouts[lo:hi].contiguous(), weights_phases, weights_sources[lo:hi].contiguous(),
                        out_of_bounds)
    else:
        local = torch.zeros((0, N), dtype=torch.float32, device=features.device)
    if not gather:
        return local, (lo, hi)
    if world == 1:
        return local
    rows = (K + world - 1) // world  # equal-sized slabs for the collective: short shards are zero padded
    slab = torch.zeros((rows, N), dtype=torch.float32, device=features.device)
    slab[: hi - lo] = local
    full = torch.empty((world * rows, N), dtype=torch.float32, device=features.device)
    dist.all_gather_into_tensor(full.view(-1), slab.view(-1), group=group)
    out = torch.empty((K, N), dtype=torch.float32, device=features.device)
    for r in range(world):
        a, b = shard_bounds(K, r, world)
        out[a:b] = full[r * rows: r * rows + (b - a)]
    return out


class PeerSlabGather:
    """All-gather of equally sized float32 slabs through NVLink/NVSwitch peer memory, moved by the copy
    engines instead of SM kernels.

    The matched-filter kernel is persistent and owns every SM (one CTA each, all of its shared memory), so
    an NCCL kernel enqueued beside it only runs in its tail.  Here every rank PUSHES its finished slab
    straight into the gather buffers of its peers (torch symmetric memory gives the peer mappings;
    ``cudaMemcpyAsync`` on a side stream runs on the DMA engines) while the next template batch computes; a
    device-side barrier per batch tells every rank that the batch is complete everywhere.
    """

    def __init__(self, n_elements_per_rank, device, group=None, n_buffers=2, n_streams=None):
        import os

        import torch.distributed._symmetric_memory as symm_mem

        if n_streams is None:  # one copy stream per peer keeps more copy engines / NVLink ports busy at once
            n_streams = int(os.environ.get("BPMF_GATHER_STREAMS", "0")) or 7

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
