"""Launched by torchrun on >= 2 GPUs (tests/test_gpu_parity.py::test_stream_sharding_nccl, or by hand):
frame-range sharding of ONE long stream over NCCL, device-resident, against the unsharded library
result and the true signal.  config-4 geometry (nfft=4096 hop=1024) on a shortened stream."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from vv_dsp_b200 import Stft, sharding
    for nfft, hop, n in ((4096, 1024, 48000 * 60), (2048, 512, 48000 * 20 + 333)):
        g = torch.Generator(device=dev).manual_seed(99)          # same stream on every rank
        x = torch.rand(n, device=dev, generator=g) * 2 - 1
        with Stft(nfft, hop, "hann") as h:
            h.set_stream(torch.cuda.current_stream().cuda_stream)
            s0, s1 = sharding.owned_samples(n, nfft, hop, world, rank)
            spec = sharding.stream_stft(h, x[s0:s1].contiguous(), n)
            whole = h.batch_forward(x[None], "complex", "valid")[0]
            frames = whole.shape[0]
            f0, f1 = sharding.frame_range(frames, world, rank)
            torch.cuda.synchronize()
            assert torch.equal(spec, whole[f0:f1]), "sharded STFT differs from the unsharded one"
            w = torch.hann_window(nfft, periodic=False, device=dev)
            y = sharding.stream_istft(h, spec, n, w)
            ywhole = h.batch_inverse(whole[None], n, True)[0]
            torch.cuda.synchronize()
            lo, hi = max(s0, nfft), min(s1, n - nfft)
            d = (y[lo - s0: hi - s0] - ywhole[lo:hi]).double()
            rel = float(torch.linalg.vector_norm(d) / torch.linalg.vector_norm(ywhole[lo:hi].double()))
            rt = float(torch.linalg.vector_norm((y[lo - s0: hi - s0] - x[lo:hi]).double()) / torch.linalg.vector_norm(x[lo:hi].double()))
            assert rel < 2e-6, rel
            assert rt < 1e-5, rt
            # the halo scheme of the C library (vv_dsp_stft_shard_inverse): bit-identical to the unsharded call
            hs = sharding.stream_stft_halo(h, x[s0:s1].contiguous(), n)
            hf = nfft // hop - 1 if rank else 0
            yh = sharding.stream_istft_halo(h, hs, n)
            torch.cuda.synchronize()
            assert torch.equal(hs[hf:], whole[f0:f1]) and torch.equal(yh, ywhole[s0:s1]), "halo-scheme shard differs from the unsharded call"
            if rank == 0:
                print(f"nccl stream sharding ok: nfft={nfft} hop={hop} n={n} world={world} frames={frames} "
                      f"vs-unsharded {rel:.2e} round-trip {rt:.2e}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
