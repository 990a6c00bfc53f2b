"""Multi-GPU parity check, run under torchrun (one rank per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
      --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_parity.py

Every rank assembles its partition on its GPU through the C ABI (continuity,
momentum-UVW, nodal gradients), the library exchanges the shared rows / shared
nodes over NCCL (nw_linsys_load_complete, nw_nodal_grad_edge), and the owned
rows of every rank are compared entry by entry (rel. 1e-12, cancellation-aware
scale) with the CPU oracle's serial assembly of the whole mesh.  Test
infrastructure: the oracle is only the checker.  tests/test_gpu_parity.py
launches this when two GPUs are visible.

  NW_MGPU_MESH=hybrid8  (world size 8): the reference's own decomposition
      reg_tests/mesh/hybrid.g.8.0-7 (tet / pyramid / hex, shared nodes with up
      to 4 sharers, neighbour sets that differ per exchange object) against the
      serial oracle on the union mesh.
  NW_MGPU_DELAY=1: afterwards, rank 1 arrives late (longer than
      NW_P2P_TIMEOUT_S) at one nodal sum; the waiting ranks must report
      NW_ERR_COMM and leave the field untouched instead of adding stale data."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.dirname(HERE), os.path.join(os.path.dirname(HERE), "oracle"), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

import oracle_py as orc  # noqa: E402
import parity_util as pu  # noqa: E402


def delay_test(P, ctx, mesh, rank, b):
    """a neighbour that is late by more than the timeout must
    produce NW_ERR_COMM, never a silent add of stale window contents.  Rank 1
    sleeps before its push; every rank that waits for it must (a) get the error
    at its next synchronisation and (b) find the field it exchanged equal to
    its own partial sums (nothing accumulated)."""
    import time
    mesh.register("delay_probe", P.NW_NODE, 1)
    mine = np.full(b.n_nodes, float(rank + 1))
    mesh.upload("delay_probe", mine)
    ctx.sync()
    dist.barrier()
    if rank == 1:
        time.sleep(float(os.environ.get("NW_P2P_TIMEOUT_S", "2")) + 2.0)
    err = None
    try:
        mesh.parallel_sum("delay_probe")
        ctx.sync()
    except P.NwError as e:
        err = str(e)
    got = None
    try:
        got = mesh.download("delay_probe")
    except P.NwError as e:
        err = err or str(e)
    if rank == 1 or err is None:
        # the late rank finds its peers' pushes waiting; a rank whose flag set
        # does not contain rank 1 is not affected
        ok = True
    else:
        ok = "did not arrive" in err and (got is None or np.array_equal(got, mine))
    return {"rank": rank, "error": err, "ok": bool(ok)}


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    dims = tuple(int(x) for x in os.environ.get("NW_MGPU_DIMS", "14,12,20").split(","))
    periodic = os.environ.get("NW_MGPU_PERIODIC", "0") == "1"
    lengths = (50.0, 40.0, 30.0)
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    P = pu.pkg()
    ctx = P.Context(local)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid = torch.tensor(list(P.Context.comm_unique_id()), dtype=torch.uint8,
                           device="cuda")
    dist.broadcast(uid, 0)
    ctx.comm_init(bytes(uid.cpu().tolist()), world, rank)

    which = os.environ.get("NW_MGPU_MESH", "box")
    if which == "hybrid8":
        assert world == 8, "the hybrid.g.8 decomposition needs 8 ranks"
        case = pu.DecomposedRealMesh(rank)
        full = pu.DecomposedRealMesh(None)
        dims, periodic = "hybrid.g.8", False
    else:
        kw = dict(dims=dims, lengths=lengths, periodic=(periodic, periodic))
        case = pu.Case(nranks=world, rank=rank, **kw)
        full = pu.Case(**kw)
    b = case.box
    mesh = b.make_mesh(ctx, tile_nodes=48)
    pu.upload_state(P, mesh, case)

    # hid -> gid of every rank (columns refer to nodes other ranks own)
    mine_map = np.array([[int(b.own_hid[l]), int(b.gid[l])]
                         for l in range(b.n_nodes)], dtype=np.int64)
    allmaps = [None] * world
    dist.all_gather_object(allmaps, mine_map)
    hid2gid = {}
    for m in allmaps:
        for h, g in m:
            hid2gid[int(h)] = int(g)
    gid2ser = {int(full.box.gid[l]): int(full.box.own_hid[l])
               for l in range(full.box.n_nodes)}
    ser = lambda h: gid2ser[hid2gid[int(h)]]

    gfull = full.oracle_graph()
    fmdot = full.oracle_mdot()
    fpec = full.oracle_pecfac(orc.peclet("classic", 1.0))
    res = {}
    plain = {}  # worst |got - ref| / |ref| (no cancellation-aware scale)

    def check_system(name, ls, oref, nrhs):
        vals, rhs = ls.values()
        gr = ls.graph()
        s = ls.sizes
        erows, ecols = ls.extra()
        fv, fr = oref.get()
        fa, fra = oref.get_abs()
        fdict, adict = {}, {}
        for r in range(gfull.num_rows_owned):
            for k in range(gfull.row_start_owned[r], gfull.row_start_owned[r + 1]):
                fdict[(r, int(gfull.cols[k]))] = fv[k]
                adict[(r, int(gfull.cols[k]))] = fa[k]
        rows_arr = np.concatenate([gr["rows"][:s.num_nonzeros_owned], erows])
        cols_arr = np.concatenate([gr["cols"][:s.num_nonzeros_owned], ecols])
        vv = np.concatenate([vals[:s.num_nonzeros_owned],
                             vals[s.num_nonzeros_owned + s.num_nonzeros_shared:]])
        periodic_rows = set(gr["periodic_rows"].tolist())
        worst, n, wplain = 0.0, 0, 0.0
        seen = set()
        for r_, c_, v_ in zip(rows_arr, cols_arr, vv):
            if int(r_) in periodic_rows:
                continue
            key = (ser(r_), ser(c_))
            assert key in fdict and key not in seen, key
            seen.add(key)
            worst = max(worst, abs(v_ - fdict[key]) /
                        (pu.TOL * max(adict[key], abs(fdict[key]), 1e-300)))
            if fdict[key] != 0.0:
                wplain = max(wplain, abs(v_ - fdict[key]) / abs(fdict[key]))
            n += 1
        mine_ser = {ser(s.i_lower + i) for i in range(s.num_rows_owned)
                    if s.i_lower + i not in periodic_rows}
        expected = sum(1 for (r, c) in fdict if r in mine_ser)
        assert n == expected, (name, n, expected)
        rhs = rhs.reshape(nrhs, -1)
        for i in range(s.num_rows_owned):
            if s.i_lower + i in periodic_rows:
                continue
            sr = ser(s.i_lower + i)
            for d in range(nrhs):
                worst = max(worst, abs(rhs[d, i] - fr[d, sr]) /
                            (pu.TOL * max(fra[d, sr], abs(fr[d, sr]), 1e-300)))
                if fr[d, sr] != 0.0:
                    wplain = max(wplain, abs(rhs[d, i] - fr[d, sr]) / abs(fr[d, sr]))
        res[name] = worst
        plain[name] = wplain

    # mdot on this rank's edges must equal the serial value of the same edge
    mesh.mdot_edge(1.0, 1.0)
    mdot = mesh.download("mass_flow_rate")
    key_full = {(int(full.box.gid[a]), int(full.box.gid[c_])): i
                for i, (a, c_) in enumerate(full.edges.reshape(-1, 2))}
    idx = np.array([key_full[(int(b.gid[a]), int(b.gid[c_]))]
                    for a, c_ in case.edges.reshape(-1, 2)])
    res["mdot"] = pu.scaled_err(mdot, fmdot[idx],
                                np.abs(fmdot[idx]) + 1e-3 * np.max(np.abs(fmdot)))
    mesh.upload("mass_flow_rate", fmdot[idx])
    mesh.upload("peclet_factor", fpec[idx])

    ls = P.LinearSystem(mesh, P.NW_LINSYS_HYPRE, 1)
    ls.buildEdgeToNodeGraph()
    ls.finalizeLinearSystem()
    ls.zeroSystem()
    ls.assemble_continuity_edge(**pu.CONT_OPTS)
    ls.loadComplete()
    ocont = pu.oracle_continuity(full, gfull)
    check_system("continuity", ls, ocont, 1)
    # residual norm over all ranks (ncclAllReduce) == serial norm
    n2 = ls.rhs_norm2_global()
    fr, fra = ocont.get()[1], ocont.get_abs()[1]
    pr = set(gfull.periodic_rows.tolist())
    keep = np.array([r not in pr for r in range(gfull.num_rows_owned)])
    ref2 = float(np.sum(fr[0, :gfull.num_rows_owned][keep] ** 2))
    res["norm2_global"] = abs(n2[0] - ref2) / (1e-10 * ref2)
    # eager exchange (push fused into the assembly kernel): bit-
    # identical to the default order; a reader completes the exchange, a
    # writer is refused while the shared rows travel
    eager = {}
    v0, r0 = ls.values()
    ls.set_eager_exchange(True)
    ls.zeroSystem()
    ls.assemble_continuity_edge(**pu.CONT_OPTS)
    ls.loadComplete()
    v1, r1 = ls.values()
    eager["continuity_bit_identical"] = bool(
        np.array_equal(v0, v1) and np.array_equal(r0, r1))
    ls.zeroSystem()
    ls.assemble_continuity_edge(**pu.CONT_OPTS)
    refused = False
    if ls.halo_transport() == "peer_memory":
        try:
            ls.assemble_continuity_edge(**pu.CONT_OPTS)
        except Exception:
            refused = True
    else:
        refused = None
    v2, r2 = ls.values()  # no loadComplete: the read completes the exchange
    ls.loadComplete()     # ... and this is then a no-op
    v3, r3 = ls.values()
    # (NCCL transport: no eager exchange, the read sees the un-summed rows)
    eager["reader_completes"] = bool(
        (refused is None or
         (np.array_equal(v0, v2) and np.array_equal(r0, r2))) and
        np.array_equal(v0, v3) and np.array_equal(r0, r3))
    eager["writer_refused"] = refused
    ls.set_eager_exchange(False)
    ls.close()

    ls = P.LinearSystem(mesh, P.NW_LINSYS_HYPRE_UVW, 3)
    ls.buildEdgeToNodeGraph()
    ls.finalizeLinearSystem()
    ls.zeroSystem()
    ls.assemble_momentum_edge("viscosity", **pu.MOM_OPTS)
    ls.loadComplete()
    check_system("momentum_uvw", ls,
                 pu.oracle_momentum(full, gfull, fmdot, fpec, uvw=True), 3)
    linsys_transport = ls.halo_transport()
    v0, r0 = ls.values()
    ls.set_eager_exchange(True)
    ls.zeroSystem()
    ls.assemble_momentum_edge("viscosity", **pu.MOM_OPTS)
    ls.loadComplete()
    v1, r1 = ls.values()
    eager["momentum_bit_identical"] = bool(
        np.array_equal(v0, v1) and np.array_equal(r0, r1))
    ls.close()

    # nodal gradients incl. the shared-node sum over NCCL
    gid2loc_full = {int(g): l for l, g in enumerate(full.box.gid)}
    for phi, d1 in (("pressure", 1), ("velocity", 3)):
        mesh.register("g_" + phi, P.NW_NODE, d1 * 3)
        os.environ["NW_HALO_OVERLAP"] = "0"  # read per call
        mesh.nodal_grad_edge(phi, "g_" + phi)
        plain_order = mesh.download("g_" + phi).copy()
        del os.environ["NW_HALO_OVERLAP"]
        mesh.nodal_grad_edge(phi, "g_" + phi)  # push fused into the kernel
        got = mesh.download("g_" + phi).reshape(b.n_nodes, d1 * 3)
        eager["grad_%s_bit_identical" % phi] = bool(
            np.array_equal(plain_order.reshape(got.shape), got))
        ref = orc.nodal_grad_edge(d1, 3, full.edges, full.fields[phi], full.area,
                                  full.fields["dual_nodal_volume"], full.n_nodes)
        ref = ref.reshape(full.n_nodes, d1 * 3)
        if periodic:  # post_work: periodic_field_update, slaves included
            ref = pu.periodic_field_update(ref, full.box.hid, full.box.own_hid)
        scale = 1e-12 * (np.max(np.abs(ref)) + 1e-300)
        worst = 0.0
        for l in range(b.n_nodes):  # owned AND shared copies carry the total
            worst = max(worst, float(np.max(np.abs(
                got[l] - ref[gid2loc_full[int(b.gid[l])]])) / scale))
        res["grad_" + phi] = worst

    # the pair gradient travels as one 6-component exchange: same bits as two
    # single calls
    mesh.register("g_density", P.NW_NODE, 3)
    mesh.register("g_pair_a", P.NW_NODE, 3)
    mesh.register("g_pair_b", P.NW_NODE, 3)
    mesh.nodal_grad_edge("density", "g_density")
    mesh.nodal_grad_edge_pair("pressure", "g_pair_a", "density", "g_pair_b")
    eager["grad_pair_bit_identical"] = bool(
        np.array_equal(mesh.download("g_pair_a"), mesh.download("g_pressure")) and
        np.array_equal(mesh.download("g_pair_b"), mesh.download("g_density")))

    # copy_owned_to_shared: poison the non-owned copies, every copy must come
    # back equal to the serial field (bit-exact: a pure copy)
    lo, hi = int(b.offsets[rank]), int(b.offsets[rank + 1]) - 1
    mine = (b.own_hid >= lo) & (b.own_hid <= hi)
    for name, nc in (("pressure", 1), ("velocity", 3)):
        ref = full.fields[name].reshape(full.n_nodes, nc)
        loc = np.array([ref[gid2loc_full[int(g)]] for g in b.gid])
        poisoned = loc.copy()
        poisoned[~mine] = -777.0
        mesh.upload(name, poisoned.reshape(case.fields[name].shape))
        mesh.copy_owned_to_shared(name)
        got = mesh.download(name).reshape(b.n_nodes, nc)
        res["copy_" + name] = 0.0 if np.array_equal(got, loc) else float("inf")

    peer_memory = ctx.peer_memory()
    nodal_transport = mesh.halo_transport()
    delay = None
    if os.environ.get("NW_MGPU_DELAY") == "1" and ctx.peer_memory():
        delay = delay_test(P, ctx, mesh, rank, b)
        ctx = None  # the context is unusable after a communication error

    allres = [None] * world
    dist.all_gather_object(allres, res)
    allplain = [None] * world
    dist.all_gather_object(allplain, plain)
    alldelay = [None] * world
    dist.all_gather_object(alldelay, delay)
    alleager = [None] * world
    dist.all_gather_object(alleager, eager)
    if rank == 0:
        worst = max(max(r.values()) for r in allres)
        out = {"world": world, "mesh": which, "dims": dims, "periodic": periodic,
               "peer_memory": peer_memory, "nodal_transport": nodal_transport,
               "linsys_transport": linsys_transport,
               "tolerance": "scaled: |got-ref| <= 1e-12 max(|ref|, sum |contributions|)",
               "worst_scaled_error": worst,
               "eager_exchange": alleager,
               "worst_plain_relative_error": {
                   k: max(p_[k] for p_ in allplain) for k in allplain[0]},
               "per_rank": allres}
        if delay is not None:
            out["delayed_rank_test"] = alldelay
        print(json.dumps(out))
        assert worst < 1.0, allres
        assert all(v is not False for e in alleager for v in e.values()), alleager
        if delay is not None:
            assert all(d["ok"] for d in alldelay), alldelay
            assert any(d["error"] for d in alldelay if d["rank"] != 1), alldelay
    if ctx is None:
        dist.destroy_process_group()
        os._exit(0)  # handles of a failed context are not torn down in order
    mesh.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
