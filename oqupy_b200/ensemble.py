"""Ensembles of independent runs sharded over the GPUs of one box.

A single TEMPO / PT-TEMPO run is time-sequential and stays on one GPU.  Independent
parameter points (couplings, temperatures, control perturbations) are assigned
cyclically to ranks (``index % world_size``: balances the chi(alpha, T) cost gradient
of a parameter grid); no data-path collective exists.  The ONLY collective is the
final gather of the per-member dynamics (``all_gather_into_tensor``: NCCL over NVLink
on GPUs, gloo in the CPU tests).
"""
import os

import numpy as np
import torch
import torch.distributed as dist


def shard_indices(n_members, rank, world_size):
    """Members owned by ``rank``: rank, rank + W, rank + 2W, ..."""
    return list(range(rank, n_members, world_size))


def _run_concurrent(mine, run_member, concurrent, make_ops, background=False,
                    thread_init=None):
    """This rank's members on ``concurrent`` host threads, each with its own ops object and
    (on a GPU) its own CUDA stream: the C-ABI calls release the GIL, a TEMPO step is one C
    call, and the small cooperative SVD launches of different members overlap on the SMs
    (measured: 74 -> 832 aggregate TEMPO steps/s on one B200 with 16 members in flight,
    profiles/r01_ensemble_one_gpu_v10.jsonl).  Results are bit-identical to serial runs.
    ``background=True`` returns at once with a ``join()`` callable that waits for the threads
    and hands back the results."""
    import queue  # pylint: disable=import-outside-toplevel
    import threading  # pylint: disable=import-outside-toplevel
    todo = queue.Queue()
    for k, i in enumerate(mine):
        todo.put((k, i))
    results, errors = [None] * len(mine), []

    def worker():
        try:
            ops = make_ops() if make_ops is not None else None
            if thread_init is not None:
                thread_init(ops)
            on_gpu = ops is not None and getattr(ops, "name", "") == "cuda"
            # background work (members re-run next to a lock-step batch that fills every SM)
            # goes on a HIGH-PRIORITY stream: its short cooperative launches take the SMs that
            # free up first instead of queueing behind the batch's remaining CTAs
            stream = torch.cuda.Stream(device=ops.device, priority=-1 if background else 0) \
                if on_gpu else None
            while True:
                try:
                    k, i = todo.get_nowait()
                except queue.Empty:
                    return
                if stream is not None:
                    with torch.cuda.stream(stream):
                        res = run_member(i, ops)
                        stream.synchronize()
                else:
                    res = run_member(i, ops)
                results[k] = np.asarray(res)
        except Exception as exc:  # pylint: disable=broad-except
            errors.append(exc)

    threads = [threading.Thread(target=worker) for _ in range(min(concurrent, len(mine)))]
    for th in threads:
        th.start()

    def join():
        for th in threads:
            th.join()
        if errors:
            raise errors[0]
        return results
    join.alive = lambda: any(th.is_alive() for th in threads)
    return join if background else join()


def run_ensemble(n_members, run_member, device=None, group=None, concurrent=1,
                 make_ops=None, run_batch=None, timings=None):
    """Run ``run_member(i)`` (-> real or complex ndarray, same shape for all i) for the
    members of this rank and gather everything on every rank.

    ``run_batch(indices) -> sequence of per-member arrays`` (instead of ``run_member``) hands
    this rank's whole share to one call: the lock-step engine (:func:`tempo_grid`) steps all of
    them with one kernel launch per time step.

    ``concurrent`` > 1 keeps that many members of this rank in flight at once (one host
    thread + ops object + CUDA stream each); ``run_member`` is then called as
    ``run_member(i, ops)`` with ``ops = make_ops()`` created once per thread (e.g.
    ``lambda: CudaOps(local_rank)``).

    Returns an ndarray (n_members, *member_shape)."""
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    import time  # pylint: disable=import-outside-toplevel
    t_run = time.perf_counter()
    mine = shard_indices(n_members, rank, world)
    if run_batch is not None:
        results = [np.asarray(r) for r in run_batch(mine)]
        assert len(results) == len(mine)
    elif concurrent > 1:
        results = _run_concurrent(mine, run_member, concurrent, make_ops)
    else:
        results = [np.asarray(run_member(i)) for i in mine]
    if timings is not None:
        timings["members_s"] = time.perf_counter() - t_run
    if world == 1:
        return np.stack(results) if results else np.zeros((0,))
    t_gather = time.perf_counter()
    # every rank learns shape/dtype from rank 0 (which always owns member 0)
    meta = [None]
    if rank == 0:
        meta = [(results[0].shape, results[0].dtype.str)]
    dist.broadcast_object_list(meta, src=0, group=group)
    shape, dtype = meta[0]
    dtype = np.dtype(dtype)
    per_rank = (n_members + world - 1) // world
    buf = np.zeros((per_rank,) + tuple(shape), dtype=dtype)
    for k, r in enumerate(results):
        buf[k] = r
    is_complex = np.iscomplexobj(buf)
    flat = np.ascontiguousarray(buf).view(np.float64) if is_complex else \
        np.ascontiguousarray(buf, dtype=np.float64)
    send = torch.from_numpy(flat.reshape(-1).copy())
    if device is not None:
        send = send.to(device)
    recv = torch.empty(world * send.numel(), dtype=send.dtype, device=send.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    allr = recv.cpu().numpy().reshape((world, per_rank) + flat.shape[1:])
    if is_complex:
        allr = allr.view(np.complex128).reshape((world, per_rank) + tuple(shape))
    out = np.zeros((n_members,) + tuple(shape), dtype=dtype)
    for r in range(world):
        for k, i in enumerate(shard_indices(n_members, r, world)):
            out[i] = allr[r, k]
    if timings is not None:
        timings["gather_s"] = time.perf_counter() - t_gather
    return out


def tempo_member(influences, propagators, initial_state, dkmax, epsrel, num_steps,
                 unitary=None, ops=None):
    """One TEMPO run on this rank's GPU; returns states (num_steps+1, d, d)."""
    from .backends import TempoBackend  # pylint: disable=import-outside-toplevel
    infl = np.asarray(influences)
    rho0 = np.asarray(initial_state, dtype=np.complex128)
    d = rho0.shape[0]
    d2 = d * d

    def influence(dk):
        return None if dk < 0 else infl[dk]

    be = TempoBackend(rho0, influence, np.eye(d) if unitary is None else unitary,
                      propagators, np.ones(d2), np.ones(d2), dkmax, epsrel, ops=ops)
    _, s0 = be.initialize()
    states = [s0]
    for _ in range(num_steps):
        states.append(be.compute_step()[1])
    return np.array(states).reshape(-1, d, d)


def broadcast_process_tensor(pt, src=0, device=None, group=None, ops=None):
    """Share ONE built process tensor with every rank ("build on GPU 0 and
    ncclBroadcast"): rank ``src`` passes its :class:`DeviceProcessTensor`, the other ranks
    pass ``None`` and get a device-resident copy.  Sites and caps travel as device tensors
    (NCCL over NVLink on GPUs -- no host staging; gloo in the CPU tests); only the shapes go
    through ``broadcast_object_list``.  Ensembles that reuse one bath over many system
    Hamiltonians (BASELINE configs[2]) then run ``dynamics_device`` on every GPU."""
    from .process_tensor import DeviceProcessTensor  # pylint: disable=import-outside-toplevel
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return pt
    rank = dist.get_rank(group)
    meta = [None]
    if rank == src:
        meta = [{"d": pt.hilbert_space_dimension, "dt": pt.dt,
                 "tin": pt.transform_in, "tout": pt.transform_out,
                 "sites": [tuple(int(x) for x in pt.get_mpo_tensor_device(k).shape)
                           for k in range(len(pt))],
                 "caps": [int(pt.get_cap_tensor_device(k).numel())
                          for k in range(len(pt) + 1)]}]
    dist.broadcast_object_list(meta, src=src, group=group)
    m = meta[0]
    if rank != src:
        pt = DeviceProcessTensor(m["d"], dt=m["dt"], transform_in=m["tin"],
                                 transform_out=m["tout"], ops=ops)
    dev = device if device is not None else getattr(pt._ops, "device", None)  # pylint: disable=protected-access
    caps = []
    for k, shape in enumerate(m["sites"]):
        if rank == src:
            t = pt.get_mpo_tensor_device(k).contiguous()
        else:
            t = torch.empty(shape, dtype=torch.complex128, device=dev)
        # complex tensors travel as their real view (NCCL has no complex dtype)
        dist.broadcast(torch.view_as_real(t), src=src, group=group)
        if rank != src:
            pt.set_mpo_tensor_device(k, t)
    for k, n in enumerate(m["caps"]):
        if rank == src:
            c = pt.get_cap_tensor_device(k).contiguous()
        else:
            c = torch.empty(n, dtype=torch.complex128, device=dev)
        dist.broadcast(torch.view_as_real(c), src=src, group=group)
        caps.append(c)
    if rank != src:
        pt._caps = caps  # pylint: disable=protected-access
    return pt


def tempo_grid(influences, initial_state, unitary, propagators, dkmax, epsrel, num_steps,
               device=None, group=None, ops=None, chunk=4096, chi_cap=None,
               fallback_concurrency=12, timings=None, check_every=10, reserve=24):
    """BASELINE configs[4]: an ensemble of independent TEMPO runs (one per parameter point)
    sharded over the ranks and, on every rank, advanced in LOCK-STEP by the batched engine
    (:class:`oqupy_b200.batch.BatchedTempoBackend`: one kernel launch per time step for all
    members of the rank).  ``influences``: (n_members, dkmax+1, d2, d2) influence matrices by
    member and dk (what each member's ``influence(dk)`` callback would return,
    oqupy/tempo.py:969-1020).  A member whose bond dimension outgrows the shared-memory
    resident SVD (chi * d2 > 128, or an intermediate operand beyond 14272 elements) is re-run alone on the general device backend
    (:func:`tempo_member`) -- still on the GPU.  Returns the dynamics
    (n_members, num_steps+1, d, d) on every rank (one gather at the end), and the indices of
    the members that took the general path."""
    from .batch import BatchedTempoBackend  # pylint: disable=import-outside-toplevel
    infl = np.asarray(influences)
    n_members = infl.shape[0]
    rho0 = np.asarray(initial_state, dtype=np.complex128)
    d = rho0.shape[-1]
    d2 = d * d
    rerun = []

    def run_batch(indices):
        out = []
        for c0 in range(0, len(indices), chunk):
            idx = list(indices[c0:c0 + chunk])
            st0 = np.broadcast_to(rho0.reshape(-1, d2), (len(idx), d2)) if rho0.ndim == 2 \
                else rho0[idx].reshape(len(idx), d2)
            import time  # pylint: disable=import-outside-toplevel
            t_0 = time.perf_counter()
            be = BatchedTempoBackend(st0, infl[idx], unitary, propagators, np.ones(d2),
                                     np.ones(d2), dkmax, epsrel, chi_cap=chi_cap, ops=ops)
            _, s0 = be.initialize()
            if timings is not None:
                timings["setup_s"] = timings.get("setup_s", 0.0) + time.perf_counter() - t_0
                t_0 = time.perf_counter()
            cuda = getattr(ops, "name", "cuda") == "cuda"
            dev_index = ops.device.index if (cuda and ops is not None) else 0

            def member(k, member_ops=None):
                st = rho0 if rho0.ndim == 2 else rho0[idx[k]]

                def member_props(step, pos_=k):
                    return tuple(np.asarray(x)[pos_] if np.asarray(x).ndim == 3 else x
                                 for x in propagators(step))
                return tempo_member(infl[idx[k]], member_props, st, dkmax, epsrel,
                                    num_steps, unitary=unitary,
                                    ops=member_ops if member_ops is not None else ops)

            def overflowed(done):
                status = be.info()["status"]
                over = []
                for k in np.nonzero(status)[0]:
                    if int(status[k]) != 2:
                        raise RuntimeError(
                            f"lock-step TEMPO member {idx[k]}: status {int(status[k])}")
                    if int(k) not in done:
                        over.append(int(k))
                return over

            def start(over):
                """The general device path for members that outgrew shared memory: several
                in flight at once, each on its own host thread and stream, WHILE the
                lock-step engine carries on with the others."""
                if cuda:
                    from ._lib import CudaOps  # pylint: disable=import-outside-toplevel
                    # the persistent batch launch leaves `reserve` SMs free from now on, and the
                    # re-runs keep their cooperative SVD grids within them (<= 5 pair slots x
                    # 4 slices for operands up to 160 columns)
                    be.reserve_sms(reserve)
                    return _run_concurrent(over, member, min(fallback_concurrency, len(over)),
                                           lambda: CudaOps(dev_index), background=True,
                                           thread_init=lambda o: o.svd_config("max_slices", 4))
                res = [member(k) for k in over]
                join_ = lambda: res      # noqa: E731
                join_.alive = lambda: False
                return join_

            # bond dimensions saturate within a few memory times: look for members that left
            # the lock-step path after 3 dkmax steps and start their re-runs right away
            probe = min(num_steps, 3 * dkmax)
            first = min(probe, max(2, dkmax // 2))
            parts = [s0[None], be.compute_steps(first, strict=False)]
            lpt = os.environ.get("B200_BATCH_REBALANCE", "1") != "0"
            if lpt:
                be.rebalance()       # heavy members (large bond dimensions) first from now on
            if probe > first:
                parts.append(be.compute_steps(probe - first, strict=False))
            if lpt:
                be.rebalance()
            redo = {}
            pending = []
            early = overflowed(redo)
            if early:
                for k in early:
                    redo[k] = None
                pending.append((early, start(early)))
            # the rest in chunks: a member that leaves the lock-step path late is noticed
            # within `check_every` steps and re-run next to the batch instead of after it
            done_steps = probe
            while done_steps < num_steps:
                nxt = min(check_every, num_steps - done_steps)
                parts.append(be.compute_steps(nxt, strict=False))
                done_steps += nxt
                if pending and cuda and not any(j.alive() for _, j in pending):
                    be.reserve_sms(0)
                if done_steps < num_steps:
                    more = overflowed(redo)
                    if more:
                        for k in more:
                            redo[k] = None
                        pending.append((more, start(more)))
            states = np.concatenate(parts)
            states = np.swapaxes(states, 0, 1).reshape(len(idx), num_steps + 1, d, d).copy()
            if timings is not None:
                timings["steps_s"] = timings.get("steps_s", 0.0) + time.perf_counter() - t_0
                t_0 = time.perf_counter()
            late = overflowed(redo)
            if late:
                for k in late:
                    redo[k] = None
                pending.append((late, start(late)))
            for over, join in pending:
                for k, res in zip(over, join()):
                    states[k] = res
            rerun.extend(int(idx[k]) for k in redo)
            if timings is not None:
                timings["general_path_wait_s"] = (timings.get("general_path_wait_s", 0.0)
                                                  + time.perf_counter() - t_0)
            out.extend(states)
        return out

    res = run_ensemble(n_members, None, device=device, group=group, run_batch=run_batch,
                       timings=timings)
    return res, rerun
