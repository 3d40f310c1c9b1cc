"""CPU tests of the N>1 path's host logic, world_size 2 over gloo: view partitioning and the
identity the multi-GPU solver relies on — the allreduced sum of the per-rank reduced-system
pieces equals the single-rank system."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from camera_calibrator_b200 import partition


def test_split_range_covers_everything_once():
    for n in (0, 1, 7, 100000):
        for world in (1, 2, 3, 8):
            cuts = [partition.split_range(n, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        partition.split_range(10, 2, 2)


def test_split_by_observations_balances_ragged_views():
    rng = np.random.default_rng(0)
    sizes = rng.integers(0, 400, 1000)
    off = np.concatenate([[0], np.cumsum(sizes)])
    cuts = [partition.split_views_by_observations(off, 4, r) for r in range(4)]
    assert cuts[0][0] == 0 and cuts[-1][1] == 1000
    assert all(cuts[i][1] == cuts[i + 1][0] for i in range(3))
    loads = [off[b] - off[a] for a, b in cuts]
    assert max(loads) - min(loads) <= 2 * 400


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from camera_calibrator_b200 import partition as part

    d = oracle.datagen_planar(11, 30, seed=3)
    _, K, q, t = oracle.estimate_init(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets)
    intr = np.array([K[0, 0], K[1, 1], K[0, 2], K[1, 2], -0.02, 0.001, 0.0005, -0.0003, 0.0002])
    a, b = part.split_range(11, world, rank)
    (u, v, X, Y, Z), off = part.slice_views([d.u, d.v, d.X, d.Y, d.Z], d.view_offsets, a, b)
    e = oracle.intrinsics_eval(u, v, X, Y, Z, off, intr, q[a:b].astype(np.float64), t[a:b].astype(np.float64))
    # the packed buffer every rank contributes: U (81) | g_u (9) | cost | n_obs
    buf = torch.tensor(np.concatenate([e["U"].ravel(), e["gu"], [e["cost"]], [float(off[-1])]]))
    dist.all_reduce(buf)
    if rank == 0:
        full = oracle.intrinsics_eval(d.u, d.v, d.X, d.Y, d.Z, d.view_offsets, intr, q.astype(np.float64),
                                      t.astype(np.float64))
        ref = np.concatenate([full["U"].ravel(), full["gu"], [full["cost"]], [float(d.view_offsets[-1])]])
        np.save(os.path.join(out_dir, "err.npy"), np.abs(buf.numpy() - ref) / np.maximum(np.abs(ref), 1e-300))
        # per-view blocks are rank-local and must be exactly the corresponding slice
        np.save(os.path.join(out_dir, "blk.npy"), np.array([np.max(np.abs(e["V"] - full["V"][a:b]))]))
    dist.destroy_process_group()


def test_allreduced_partial_systems_equal_the_full_system(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    err = np.load(tmp_path / "err.npy")
    assert err.max() <= 1e-12
    assert np.load(tmp_path / "blk.npy")[0] == 0.0


def _rig_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from camera_calibrator_b200 import partition as part
    from camera_calibrator_b200 import synth

    # the rig is partitioned by FRAME: camera extrinsics replicated, every rank contributes the
    # camera blocks U_cc, the camera gradient and the robust cost of its frames (what k_rig_decide all-reduces)
    s = synth.rig_scene(n_cams=3, n_frames=9, grid=(5, 4), seed=4, outlier_fraction=0.1)
    huber = float(np.float32(3.0) / np.float32(500.0))
    fa, fb = part.split_range(9, world, rank)
    o0, o1 = int(s.frame_offsets[fa]), int(s.frame_offsets[fb])
    e = oracle.rig_eval(s.u[o0:o1], s.v[o0:o1], s.obs_cam[o0:o1], s.obs_point[o0:o1], s.frame_offsets[fa:fb + 1] - o0,
                        s.Xw, s.q_c, s.t_c, s.q_f[fa:fb], s.t_f[fa:fb], huber)
    buf = torch.tensor(np.concatenate([e["U"].ravel(), e["gu"], [e["cost"]], [float(o1 - o0)]]))
    dist.all_reduce(buf)
    if rank == 0:
        full = oracle.rig_eval(s.u, s.v, s.obs_cam, s.obs_point, s.frame_offsets, s.Xw, s.q_c, s.t_c, s.q_f, s.t_f, huber)
        ref = np.concatenate([full["U"].ravel(), full["gu"], [full["cost"]], [float(s.frame_offsets[-1])]])
        scale = np.maximum(np.abs(ref), 1e-300)
        np.save(os.path.join(out_dir, "rig_err.npy"), np.abs(buf.numpy() - ref) / np.maximum(scale, np.abs(ref).max() * 1e-12))
        # per-frame blocks are rank-local: exactly the corresponding slice of the full evaluation
        np.save(os.path.join(out_dir, "rig_blk.npy"),
                np.array([np.max(np.abs(e["V"] - full["V"][fa:fb])), np.max(np.abs(e["W"] - full["W"][fa:fb]))]))
    dist.destroy_process_group()


def test_allreduced_partial_rig_systems_equal_the_full_system(tmp_path):
    port = _free_port()
    mp.spawn(_rig_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert np.load(tmp_path / "rig_err.npy").max() <= 1e-12
    assert np.all(np.load(tmp_path / "rig_blk.npy") == 0.0)
