"""N > 1 path on CPU: world_size-2 gloo run of pixsht.distributed.ShardedSHT (partition, pack/unpack, all-to-all) with
the host-emulation build of the kernels standing in for the GPU stages.  Results are checked against the oracle."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, lmax, res_deg, out_dir):
    for p in (ROOT, os.path.join(ROOT, "pixell.jl_b200"), HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    import pixsht
    from pixsht.transforms import PixshtLib
    from pixsht.distributed import ShardedSHT
    from helpers import synth_alm
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    lib = PixshtLib(os.path.join(HERE, "emu", "_build", "libpixsht_emu.so"))
    shape, wcs = pixsht.fullsky_geometry(res_deg * pixsht.degree)
    band = pixsht.sht_band(shape, wcs)
    sht = ShardedSHT(band, lmax, device="cpu", lib=lib)
    nc = 3
    alms = [torch.from_numpy(synth_alm(lmax, lmax, 100 + c, spin2=c > 0)) for c in range(nc)]
    a, b = sht.map_rows()
    slabs = [torch.zeros((b - a) * band.nx, dtype=torch.float64) for _ in range(nc)]
    sht.alm2map(alms, slabs)
    np.save(os.path.join(out_dir, "slab_%d.npy" % rank), np.stack([s.numpy() for s in slabs]))
    out = [torch.full((sht.nalm,), 7.0, dtype=torch.complex128) for _ in range(nc)]
    sht.map2alm(slabs, out)
    np.save(os.path.join(out_dir, "alm_%d.npy" % rank), np.stack([o.numpy() for o in out]))
    np.save(os.path.join(out_dir, "rows_%d.npy" % rank), np.array([a, b]))
    # host-resident variants (per-family pipelines through pixsht_plan_set_stage_families): same numbers as the device-tensor calls
    idx = torch.cat([torch.arange(s, e) for (s, e) in sht.alm_columns()])
    h_cols = [x.index_select(0, idx).clone() for x in alms]
    work = [torch.zeros(sht.nalm, dtype=torch.complex128) for _ in range(nc)]
    d_slabs = [torch.zeros_like(x) for x in slabs]
    h_slabs = [torch.zeros_like(x) for x in slabs]
    sht.alm2map_host(h_cols, h_slabs, work, d_slabs)
    ok = all(torch.equal(x, y) for x, y in zip(h_slabs, slabs))
    h_out = [torch.zeros_like(x) for x in h_cols]
    sht.map2alm_host(h_slabs, h_out, d_slabs, work)
    ok = ok and all(torch.allclose(x, o.index_select(0, idx), rtol=0, atol=1e-13) for x, o in zip(h_out, out))
    # two components (spin 2 alone) and one (T alone) take the other family tables
    for comps in ([1, 2], [0]):
        w2 = [torch.zeros(sht.nalm, dtype=torch.complex128) for _ in comps]
        hs2 = [torch.zeros_like(slabs[0]) for _ in comps]
        sht.alm2map_host([h_cols[c] for c in comps], hs2, w2, [torch.zeros_like(slabs[0]) for _ in comps])
        ok = ok and all(torch.equal(x, slabs[c]) for x, c in zip(hs2, comps))
    np.save(os.path.join(out_dir, "hostok_%d.npy" % rank), np.array([int(ok)]))
    dist.barrier()
    sht.close()
    # Float32 maps / complex64 alm through the same pipeline (T only)
    sht32 = ShardedSHT(band, lmax, device="cpu", lib=lib, dtype=torch.float32)
    alm32 = [alms[0].to(torch.complex64)]
    slab32 = [torch.zeros((b - a) * band.nx, dtype=torch.float32)]
    sht32.alm2map(alm32, slab32)
    out32 = [torch.zeros(sht32.nalm, dtype=torch.complex64)]
    sht32.map2alm(slab32, out32)
    np.save(os.path.join(out_dir, "slab32_%d.npy" % rank), slab32[0].numpy())
    np.save(os.path.join(out_dir, "alm32_%d.npy" % rank), out32[0].numpy())
    dist.barrier()
    sht32.close()
    dist.destroy_process_group()


def test_partitions():
    sys.path.insert(0, os.path.join(ROOT, "pixell.jl_b200"))
    from pixsht.distributed import partition_m, partition_m_weighted, partition_rings
    for mmax in (0, 1, 7, 36, 10800):
        for world in (1, 2, 3, 8):
            lists = partition_m(mmax, world)
            allm = np.sort(np.concatenate(lists))
            assert np.array_equal(allm, np.arange(mmax + 1))
            cost = [int(np.sum(mmax - l + 1)) for l in lists]
            assert max(cost) - min(cost) <= mmax + 2   # at most one (m, mmax-m) pair of imbalance
            wl = partition_m_weighted(mmax + 1.0 - np.arange(mmax + 1), world)
            assert np.array_equal(np.sort(np.concatenate(wl)), np.arange(mmax + 1))
            if mmax >= 16 * 8 * world:   # enough runs of 16 to balance: within one run of the mean
                cw = [float(np.sum(mmax + 1.0 - l)) for l in wl]
                assert max(cw) - min(cw) <= 16 * (mmax + 1)
            for l in wl:                 # runs of 16 consecutive m stay together
                assert all(np.array_equal(l[l // 16 == b], np.arange(b * 16, min(mmax + 1, b * 16 + 16))) for b in np.unique(l // 16))
            rr = partition_rings(mmax + 1, world)
            assert rr[0][0] == 0 and rr[-1][1] == mmax + 1 and all(rr[i][1] == rr[i + 1][0] for i in range(world - 1))


@pytest.mark.parametrize("world", [2])
def test_sharded_pipeline_matches_oracle(world, tmp_path):
    import subprocess
    emu_so = os.path.join(HERE, "emu", "_build", "libpixsht_emu.so")
    if not os.path.exists(emu_so):
        subprocess.check_call(["bash", os.path.join(HERE, "emu", "build_emu.sh")])
    lmax, res_deg = 36, 5.0
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(world, port, lmax, res_deg, str(tmp_path)), nprocs=world, join=True)
    import pixsht
    from helpers import synth_alm, oracle_alm2map, oracle_map2alm, rel_rms
    shape, wcs = pixsht.fullsky_geometry(res_deg * pixsht.degree)
    alms = [synth_alm(lmax, lmax, 100 + c, spin2=c > 0) for c in range(3)]
    ref = np.concatenate([oracle_alm2map(alms[0][None], shape, wcs, lmax),
                          oracle_alm2map(np.stack(alms[1:]), shape, wcs, lmax, spin=2)], axis=2)
    got = np.zeros((3, shape[1], shape[0]))
    for r in range(world):
        a, b = np.load(tmp_path / ("rows_%d.npy" % r))
        got[:, a:b, :] = np.load(tmp_path / ("slab_%d.npy" % r)).reshape(3, b - a, shape[0])
    got = got.transpose(2, 1, 0)
    assert rel_rms(got, ref) < 1e-10
    got32 = np.zeros((shape[1], shape[0]), dtype=np.float32)
    for r in range(world):
        a, b = np.load(tmp_path / ("rows_%d.npy" % r))
        got32[a:b, :] = np.load(tmp_path / ("slab32_%d.npy" % r)).reshape(b - a, shape[0])
    assert rel_rms(got32.T.astype(np.float64), ref[:, :, 0]) < 1e-5            # Float32 tolerance of north_star
    alm32 = sum(np.load(tmp_path / ("alm32_%d.npy" % r)) for r in range(world)).astype(np.complex128)
    assert rel_rms(alm32, oracle_map2alm(pixsht.Enmap(np.asfortranarray(got32.T.astype(np.float64)), wcs), lmax)[0]) < 1e-5
    assert all(int(np.load(tmp_path / ("hostok_%d.npy" % r))[0]) == 1 for r in range(world))   # host-path variants = device-tensor calls
    alm_sum = sum(np.load(tmp_path / ("alm_%d.npy" % r)) for r in range(world))
    rt = oracle_map2alm(pixsht.Enmap(ref[:, :, 0], wcs), lmax)[0]
    reb = oracle_map2alm(pixsht.Enmap(ref[:, :, 1:], wcs), lmax, spin=2)
    assert rel_rms(alm_sum[0], rt) < 1e-10 and rel_rms(alm_sum[1], reb[0]) < 1e-10 and rel_rms(alm_sum[2], reb[1]) < 1e-10
