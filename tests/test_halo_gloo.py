"""Host logic of the multi-rank path on CPU: world_size-2/3 `gloo` process groups exchange
real buffers following the library's exchange lists (the same lists the NCCL path uses on
the GPUs), for the vector halos C1-C3/C6 and the matrix ghost-row exchanges C7/C8
(source/sparse_matrix_simd.h:651-763)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import ryujin_b200 as rj

B = rj.Boundary


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _f_vec(g, c):
    return 1000.0 * c + g.astype(np.float64) + 0.25


def _f_mat(gi, gj):
    return gi.astype(np.float64) * 1.0e4 + gj.astype(np.float64) + 0.5


def _worker(rank, world, port, n_cells, q, partition=None):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        dim = len(n_cells)
        mesh = rj.CartesianMesh(n_cells, [0.0] * dim, [1.0] * dim, [B.dirichlet] * (2 * dim), partition=partition)
        host = rj.HostOfflineData(mesh, rank, world)
        od = rj.OfflineData(host, host_layout_only=True)
        neighbors, send_rows, mat_send = od.exchange_lists()
        so, rl, cols, tr = od.layout()
        g = host.global_index.astype(np.int64)
        n_owned, n_rel, ld = host.n_owned, host.n_relevant, od.ld

        # ---- vector halo: component-major vector with stride ld, k components ----
        k = 3
        v = np.full((k, ld), np.nan)
        for c in range(k):
            v[c, :n_owned] = _f_vec(g[:n_owned], c)
        reqs, recv_bufs = [], []
        for nb in neighbors:
            rows = send_rows[nb["send_offset"] : nb["send_offset"] + nb["send_count"]]
            buf = torch.from_numpy(np.ascontiguousarray(v[:, rows]))  # [component][row]
            reqs.append(dist.isend(buf, nb["rank"]))
            rb = torch.empty((k, nb["ghost_count"]), dtype=torch.float64)
            recv_bufs.append(rb)
            reqs.append(dist.irecv(rb, nb["rank"]))
        for r in reqs:
            r.wait()
        for nb, rb in zip(neighbors, recv_bufs):
            v[:, nb["ghost_start"] : nb["ghost_start"] + nb["ghost_count"]] = rb.numpy()
        for c in range(k):
            assert np.array_equal(v[c, :n_rel], _f_vec(g[:n_rel], c)), "vector halo mismatch"

        # ---- matrix ghost rows: l_ij in the sliced layout, ghost rows appended ----
        m = np.full(od.n_entries_total, np.nan)
        for i in range(n_owned):
            e = int(so[i // 32]) + i % 32 + 32 * np.arange(rl[i])
            m[e] = _f_mat(np.full(rl[i], g[i]), g[cols[e]])
        reqs, recv_bufs = [], []
        for nb in neighbors:
            ents = mat_send[nb["mat_send_offset"] : nb["mat_send_offset"] + nb["mat_send_count"]]
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(m[ents])), nb["rank"]))
            rb = torch.empty(nb["mat_recv_count"], dtype=torch.float64)
            recv_bufs.append(rb)
            reqs.append(dist.irecv(rb, nb["rank"]))
        for r in reqs:
            r.wait()
        for nb, rb in zip(neighbors, recv_bufs):
            m[nb["mat_recv_offset"] : nb["mat_recv_offset"] + nb["mat_recv_count"]] = rb.numpy()
        # every transposed slot (also those in ghost rows) now holds l_ji
        n_checked = 0
        for i in range(n_owned):
            e = int(so[i // 32]) + i % 32 + 32 * np.arange(rl[i])
            j = cols[e]
            assert np.array_equal(m[tr[e]], _f_mat(g[j], np.full(rl[i], g[i]))), "matrix ghost rows mismatch"
            n_checked += int(np.sum(j >= n_owned))
        assert (n_checked > 0) == (len(neighbors) > 0)

        # the all-reduce pattern of tau_max / restart flag (C4, C9)
        t = torch.tensor([1.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        assert t.item() == 1.0
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as exc:  # pragma: no cover
        import traceback

        q.put((rank, "error: %s\n%s" % (exc, traceback.format_exc())))


@pytest.mark.parametrize("n_cells,world,partition", [([6, 9], 2, None), ([4, 3, 7], 2, None), ([5, 11], 3, None),
                                                     ([7, 9], 4, [2, 2]), ([5, 6, 7], 4, [1, 2, 2])])
def test_exchange_lists_over_gloo(n_cells, world, partition):
    """slabs and block partitions (face, edge and corner neighbours): one buffer [component][row] per
    neighbour and exchange, as the NCCL path sends it"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_cells, q, partition)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", "rank %d: %s" % (rank, msg)


def test_device_layout_decodes_to_graph_on_host():
    """'bit-exact stencil indices' of the warp-sliced layout, checked without a GPU on the
    host-only layout (the GPU test repeats it on the uploaded arrays)."""
    for mesh in (
        rj.CartesianMesh([37, 21], [0.0, 0.0], [1.0, 1.0], [B.slip] * 4),
        rj.CartesianMesh([9, 7, 5], [0.0] * 3, [1.0] * 3, [B.dirichlet] * 6),
        rj.CartesianMesh.forward_facing_step(0),
    ):
        host = rj.HostOfflineData(mesh)
        od = rj.OfflineData(host, host_layout_only=True)
        so, rl, cols, tr = od.layout()
        assert np.array_equal(rl, np.diff(host.row_starts.astype(np.int64)))
        row_of_entry = np.full(od.n_entries_owned, -1, dtype=np.int64)
        for i in range(host.n_owned):
            e = int(so[i // 32]) + i % 32 + 32 * np.arange(rl[i])
            row_of_entry[e] = i
            r0 = int(host.row_starts[i])
            assert np.array_equal(cols[e], host.cols[r0 : r0 + rl[i]])
            assert np.array_equal(cols[tr[e]], np.full(rl[i], i))
        for i in range(host.n_owned):
            e = int(so[i // 32]) + i % 32 + 32 * np.arange(rl[i])
            assert np.array_equal(row_of_entry[tr[e]], cols[e].astype(np.int64))
        # slices are padded to their longest row
        for gsl in range(od.n_slices):
            w = (int(so[gsl + 1]) - int(so[gsl])) // 32
            assert w == int(np.max(rl[32 * gsl : 32 * gsl + 32]))
    with pytest.raises(rj.NoDeviceError):
        rj.HyperbolicModule(od)
