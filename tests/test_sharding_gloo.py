"""CPU tests (-m "not gpu"): the N>1 paths with world_size 2 over gloo.  Compute runs through the
emulator build of the library (tests/emu) because there is no GPU here; on the B200 the same
functions run with CUDA tensors over NCCL (test_gpu_parity.py / bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, HERE); sys.path.insert(0, os.path.join(HERE, "emu"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import build_emu
    from vv_dsp_b200 import Library, Stft, sharding
    from oracle.oracle import Oracle
    from _util import noise, rel_l2, spectra_close
    lib = Library(build_emu.build())
    o = Oracle()
    nfft, hop, n = case
    x = noise(4242, n)                                   # every rank can regenerate the global stream
    try:
        with Stft(nfft, hop, "hann", lib=lib) as h:
            s0, s1 = sharding.owned_samples(n, nfft, hop, world, rank)
            spec = sharding.stream_stft(h, torch.from_numpy(x[s0:s1].copy()), n)
            ref = o.stft(x, nfft, hop)
            f0, f1 = sharding.frame_range(ref.shape[0], world, rank)
            ok, frac = spectra_close(spec.numpy(), ref[f0:f1])
            assert ok, frac
            # same bits as the unsharded library call: sharding must not change the arithmetic
            whole = h.batch_forward(x[None], "complex", "valid")[0]
            assert np.array_equal(spec.numpy(), whole[f0:f1])
            w = torch.from_numpy(o.window("hann", nfft)[1].copy())
            y = sharding.stream_istft(h, spec, n, w).numpy()
            refy = o.istft(ref, nfft, hop, n)
            lo, hi = max(s0, nfft), min(s1, n - nfft)
            assert rel_l2(y[lo - s0: hi - s0], refy[lo:hi]) < 5e-5
            assert rel_l2(y[lo - s0: hi - s0], x[lo:hi]) < 1e-5
            # the precomputed plan object gives the same results as the functional form
            plan = sharding.StreamPlan(h, n, w)
            plan.x_owned.copy_(torch.from_numpy(x[s0:s1].copy()))
            for _ in range(2):
                sp = plan.stft()
                assert np.array_equal(sp.numpy(), spec.numpy())
                yp = plan.istft(sp).numpy()
                assert np.allclose(yp[lo - s0: hi - s0], y[lo - s0: hi - s0], rtol=0, atol=2e-6)
            # the halo scheme (what the C library's multi-device handle does): no partial sums exchanged, and the ranks'
            # outputs are the SAME BITS as the unsharded call
            if sharding.halo_mode(nfft, hop):
                hs = sharding.stream_stft_halo(h, torch.from_numpy(x[s0:s1].copy()), n)
                hf = nfft // hop - 1 if rank else 0
                assert hs.shape[0] == hf + (f1 - f0) and np.array_equal(hs.numpy()[hf:], whole[f0:f1])
                if rank:
                    assert np.array_equal(hs.numpy()[:hf], whole[f0 - hf:f0])
                yh = sharding.stream_istft_halo(h, hs, n).numpy()
                ywhole = h.batch_inverse(whole[None], n, True)[0]
                if nfft >= 2048:
                    assert np.array_equal(yh, ywhole[s0:s1])
                else:                                       # below 2048 the unsharded call pairs frames: equal to rounding
                    assert np.abs(yh - ywhole[s0:s1])[nfft if rank == 0 else 0:(s1 - s0) - (nfft if rank == world - 1 else 0)].max() < 2e-6
            # shard-by-signal rule: disjoint cover, results identical to the unsharded call
            B = 5
            xb = np.stack([noise(100 + i, 6000) for i in range(B)])
            b0, b1 = sharding.shard_batch(B, world, rank)
            mine = h.batch_forward(xb[b0:b1], "power", "valid")
            assert np.array_equal(mine, h.batch_forward(xb, "power", "valid")[b0:b1])
            counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
            dist.all_gather(counts, torch.tensor([b1 - b0]))
            assert sum(int(c) for c in counts) == B
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", [(256, 64, 6000), (2048, 512, 40000), (512, 200, 9001), (4096, 1024, 61001), (1024, 256, 20000)])
def test_stream_sharding_two_ranks(case, tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), case, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


def test_partition_rules():
    sys.path.insert(0, ROOT)
    from vv_dsp_b200 import sharding
    for total in (0, 1, 7, 934, 168747):
        for world in (1, 2, 3, 8):
            r = [sharding.frame_range(total, world, k) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total and all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            b = [sharding.shard_batch(total, world, k) for k in range(world)]
            assert b == r
    # config 4: 1 h @ 48 kHz, nfft=4096 hop=1024 on 8 ranks
    n, nfft, hop = 172_800_000, 4096, 1024
    spans = [sharding.owned_samples(n, nfft, hop, 8, k) for k in range(8)]
    assert spans[0][0] == 0 and spans[-1][1] == n and all(spans[i][1] == spans[i + 1][0] for i in range(7))
    assert all(s[0] % hop == 0 for s in spans)
