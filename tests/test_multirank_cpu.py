"""world_size-2/3/8 gloo tests (CPU) of the multi-rank host logic: the shared-row
and shared-node exchange lists the library builds, driven with a caller-side
transport (torch.distributed gloo) exactly as include/nalu_edge_b200.h describes
for callers without NCCL.  The per-rank numbers come from the CPU oracle; what is
under test is that the product's index lists route every shared contribution to
the right slot: the partitioned assembly + halo sum must equal the serial
assembly of the whole mesh."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _exchange(send, rank, world):
    """send: {peer: int64 array}; returns {peer: array}"""
    out = {}
    for a in range(world):
        for b in range(world):
            if a == b:
                continue
            if rank == a:
                arr = send.get(b, np.zeros(0, dtype=np.int64))
                dist.send(torch.tensor([arr.size], dtype=torch.int64), b)
                if arr.size:
                    dist.send(torch.from_numpy(np.ascontiguousarray(arr)), b)
            elif rank == b:
                n = torch.zeros(1, dtype=torch.int64)
                dist.recv(n, a)
                t = torch.zeros(int(n.item()), dtype=torch.int64)
                if n.item():
                    dist.recv(t, a)
                out[a] = t.numpy()
    return out


def _exchange_f64(send, rank, world):
    out = {}
    for a in range(world):
        for b in range(world):
            if a == b:
                continue
            if rank == a:
                arr = send.get(b, np.zeros(0))
                dist.send(torch.tensor([arr.size], dtype=torch.int64), b)
                if arr.size:
                    dist.send(torch.from_numpy(np.ascontiguousarray(arr)), b)
            elif rank == b:
                n = torch.zeros(1, dtype=torch.int64)
                dist.recv(n, a)
                t = torch.zeros(int(n.item()), dtype=torch.float64)
                if n.item():
                    dist.recv(t, a)
                out[a] = t.numpy()
    return out


def _worker(rank, world, port, dims, periodic, q):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        for p in (os.path.dirname(HERE), os.path.join(os.path.dirname(HERE), "oracle"), HERE):
            if p not in sys.path:
                sys.path.insert(0, p)
        import oracle_py as orc
        import parity_util as pu
        P = pu.pkg()
        if dims == "hybrid8":
            # the reference's own 8-way decomposition (reg_tests/mesh/hybrid.g.8.*)
            case = pu.DecomposedRealMesh(rank)
        else:
            case = pu.Case(dims=dims, nranks=world, rank=rank, periodic=periodic,
                           lengths=(50.0, 40.0, 30.0))
        b = case.box
        ctx = P.Context(-1)
        mesh = b.make_mesh(ctx, tile_nodes=40)
        ls = P.LinearSystem(mesh)
        ls.buildEdgeToNodeGraph()
        ls.finalizeLinearSystem()

        # ---- shared rows: structure exchange with the caller's transport ----
        send = {}
        for peer in range(world):
            if peer == rank:
                continue
            rows, lens, cols = ls.halo_send(peer)
            if rows.size:
                send[peer] = np.concatenate([[rows.size], rows, lens, cols]).astype(np.int64)
        got = _exchange(send, rank, world)
        for peer, blob in got.items():
            if blob.size:
                n = int(blob[0])
                ls.halo_set_recv(peer, blob[1:1 + n], blob[1 + n:1 + 2 * n],
                                 blob[1 + 2 * n:])
        ls.halo_commit()

        # ---- per-rank assembly (oracle numbers in the product's layout) ----
        g = case.oracle_graph()
        o = pu.oracle_continuity(case, g)
        vals, rhs = o.get()
        mine = ls.graph()
        assert np.array_equal(mine["cols"], g.cols)
        s = ls.sizes
        erows, ecols = ls.extra()
        vals = np.concatenate([vals, np.zeros(erows.size)])
        rhs = rhs[0].copy()
        # send the tail segments in place, owner adds in ascending peer order
        vsend, rsend = {}, {}
        for peer in range(world):
            if peer == rank:
                continue
            rows, lens, cols = ls.halo_send(peer)
            if rows.size == 0:
                continue
            i0 = int(np.searchsorted(mine["row_indices_shared"], rows[0]))
            a = s.num_nonzeros_owned + mine["row_start_shared"][i0]
            vsend[peer] = vals[a:a + cols.size]
            rsend[peer] = rhs[s.num_rows_owned + i0:s.num_rows_owned + i0 + rows.size]
        vgot = _exchange_f64(vsend, rank, world)
        rgot = _exchange_f64(rsend, rank, world)
        for peer in sorted(vgot):
            vs, rr = ls.halo_recv_slots(peer)
            assert vs.size == vgot[peer].size and rr.size == rgot[peer].size
            np.add.at(vals, vs, vgot[peer])
            np.add.at(rhs, rr, rgot[peer])

        # ---- compare the owned rows with the serial assembly ----
        if dims == "hybrid8":
            full = pu.DecomposedRealMesh(None)
        else:
            full = pu.Case(dims=dims, periodic=periodic, lengths=(50.0, 40.0, 30.0))
        # serial row ids == global ids - 1 resolved; map through gid
        gfull = full.oracle_graph()
        of = pu.oracle_continuity(full, gfull)
        fv, fr = of.get()
        fa, fra = of.get_abs()
        # serial hid of a partitioned row id: via global ids
        hid2gid = {}
        for l in range(b.n_nodes):
            hid2gid[int(b.own_hid[l])] = int(b.gid[l])
        gid2ser = {int(full.box.gid[l]): int(full.box.own_hid[l]) for l in range(full.box.n_nodes)}
        # all ranks' hid->gid (columns may refer to nodes this rank does not have)
        blob = np.array([[h, gg] for h, gg in hid2gid.items()], dtype=np.int64).ravel()
        allmaps = _exchange({p: blob for p in range(world) if p != rank}, rank, world)
        for bl in allmaps.values():
            for h, gg in bl.reshape(-1, 2):
                hid2gid[int(h)] = int(gg)
        ser = lambda h: gid2ser[hid2gid[int(h)]]
        fdict, adict = {}, {}
        for r in range(gfull.num_rows_owned):
            for k in range(gfull.row_start_owned[r], gfull.row_start_owned[r + 1]):
                fdict[(r, int(gfull.cols[k]))] = fv[k]
                adict[(r, int(gfull.cols[k]))] = fa[k]
        lo = s.i_lower
        worst = 0.0
        nchecked = 0
        rows_arr = np.concatenate([mine["rows"][:s.num_nonzeros_owned], erows])
        cols_arr = np.concatenate([mine["cols"][:s.num_nonzeros_owned], ecols])
        vv = np.concatenate([vals[:s.num_nonzeros_owned], vals[s.num_nonzeros_owned + s.num_nonzeros_shared:]])
        seen = set()
        periodic_rows = set(mine["periodic_rows"].tolist())
        for r_, c_, v_ in zip(rows_arr, cols_arr, vv):
            if int(r_) in periodic_rows:
                continue
            key = (ser(r_), ser(c_))
            assert key in fdict, ("entry not in the serial matrix", key)
            assert key not in seen
            seen.add(key)
            worst = max(worst, abs(v_ - fdict[key]) / (pu.TOL * max(adict[key], 1e-300)))
            nchecked += 1
        for i in range(s.num_rows_owned):
            if lo + i in periodic_rows:
                continue
            sr = ser(lo + i)
            worst = max(worst, abs(rhs[i] - fr[0, sr]) / (pu.TOL * max(fra[0, sr], 1e-300)))
        # every serial entry of my rows must have been produced
        mine_ser_rows = {ser(lo + i) for i in range(s.num_rows_owned)
                         if lo + i not in periodic_rows}
        expected = sum(1 for (r, c) in fdict if r in mine_ser_rows)
        assert nchecked == expected, (nchecked, expected)

        # ---- shared nodes: gradient partial sums -> parallel_sum ----
        send = {}
        for peer in range(world):
            if peer != rank:
                h = mesh.halo_send(peer)
                if h.size:
                    send[peer] = h
        got = _exchange(send, rank, world)
        for peer, h in got.items():
            mesh.halo_set_recv(peer, h)
        mesh.halo_commit()
        f = case.fields
        part = orc.nodal_grad_edge(1, 3, case.edges, f["pressure"], case.area,
                                   f["dual_nodal_volume"], case.n_nodes)
        own2loc = {int(h): l for l, h in enumerate(b.own_hid)}
        gsend = {p: part[[own2loc[int(h)] for h in hs]].ravel() for p, hs in send.items()}
        ggot = _exchange_f64(gsend, rank, world)
        tot = part.copy()
        for peer in sorted(ggot):
            idx = [own2loc[int(h)] for h in got[peer]]
            np.add.at(tot, idx, ggot[peer].reshape(-1, 3))
        fullg = orc.nodal_grad_edge(1, 3, full.edges, full.fields["pressure"],
                                    full.area, full.fields["dual_nodal_volume"],
                                    full.n_nodes)
        gid2loc_full = {int(gg): l for l, gg in enumerate(full.box.gid)}
        gerr = 0.0
        for l in range(b.n_nodes):
            if b.owner[l] != rank:
                continue
            ref = fullg[gid2loc_full[int(b.gid[l])]]
            gerr = max(gerr, float(np.max(np.abs(tot[l] - ref)) /
                                   (1e-12 * (np.max(np.abs(fullg)) + 1e-300))))
        q.put((rank, "ok", worst, gerr))
    except Exception as ex:  # pragma: no cover
        import traceback
        q.put((rank, "fail", traceback.format_exc(), str(ex)))
    finally:
        try:
            dist.destroy_process_group()
        except Exception:
            pass


@pytest.mark.parametrize("world,dims,periodic", [
    (2, (5, 4, 6), (False, False)),
    (3, (4, 5, 9), (False, False)),
    (2, (5, 4, 6), (True, True)),
    (8, "hybrid8", (False, False)),
])
def test_halo_lists_route_partitioned_assembly(world, dims, periodic):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, dims, periodic, q))
             for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for r in res:
        assert r[1] == "ok", r
        assert r[2] < 1.0, r
        assert r[3] < 1.0, r
