"""CPU suite, part 3: the N > 1 host logic (batch sharding + the rollout's only collective) on the gloo backend with
world_size 2 - no GPU needed."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, T, B_global):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from diffusion_by_maxentirl_b200.dist import gather_rollout, shard_bounds, shard_noise

    g = torch.Generator().manual_seed(7)
    noise = torch.randn(T + 1, B_global, 3, 8, 8, generator=g)  # identical on every rank
    mine = shard_noise(noise, rank, world)
    lo, hi = shard_bounds(B_global, rank, world)
    assert mine.shape[1] == hi - lo and torch.equal(mine, noise[:, lo:hi])
    # stand-in "rollout": any per-sample function of the noise (the real one is per-sample independent)
    x_T = mine.sum(0).clamp(-1, 1)
    u8 = ((x_T + 1) * 127.5).clamp(0, 255).to(torch.uint8)
    energy = x_T.flatten(1).sum(1, keepdim=True)
    all_u8, all_e = gather_rollout(u8, energy)
    want_x = noise.sum(0).clamp(-1, 1)
    want_u8 = ((want_x + 1) * 127.5).clamp(0, 255).to(torch.uint8)
    assert all_u8.shape == (B_global, 3, 8, 8) and torch.equal(all_u8, want_u8)
    assert torch.allclose(all_e, want_x.flatten(1).sum(1))
    # the packed form: ONE collective for (u8 samples | fp32 energies); bit-identical payload
    from diffusion_by_maxentirl_b200.dist import PackedRollout

    pk = PackedRollout(hi - lo, (3, 8, 8), "cpu", world=world)
    pk.samples_u8.copy_(u8)
    pk.energies.copy_(energy.reshape(-1))
    p_u8, p_e = pk.all_gather()
    assert torch.equal(p_u8, want_u8) and torch.equal(p_e, all_e)
    assert pk.local.numel() == (hi - lo) * (3 * 8 * 8 + 4)
    dist.destroy_process_group()


def test_shard_and_gather_world2():
    mp.spawn(_worker, args=(2, _free_port(), 3, 6), nprocs=2, join=True)


def test_shard_bounds_cover_ragged_batches():
    from diffusion_by_maxentirl_b200.dist import shard_bounds

    for n in (0, 1, 7, 8, 513):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
