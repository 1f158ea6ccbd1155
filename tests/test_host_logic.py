"""Host-side logic that needs no GPU: synthetic inputs, the station-DB container, Tiler / partitioning."""
import os

import numpy as np
import pytest

from topowx_b200 import synth, db
from topowx_b200.interp.tiling import Tiler, partition_chunks


@pytest.fixture(scope="module")
def small_db():
    return synth.make_station_db(0, 300, synth.tile_bbox(buf=1.0), synth.Fields(), synth.make_days(1995, 1))


def test_station_db_contract(small_db, tmp_path):
    d = small_db
    assert d.var.dtype == np.float32 and d.var.shape == (365, d.stns.size)
    assert list(d.stn_ids) == sorted(d.stn_ids)                   # DB order is station-id order
    assert sum(d.mth_idx[m].size for m in range(1, 13)) == 365 and d.mth_idx[2].size == 28
    ids = d.stn_ids[[5, 2, 9]]
    obs = d.load_obs(ids, mth=3)
    assert obs.shape == (31, 3)
    assert np.array_equal(obs, d.var[d.mth_idx[3]][:, [2, 5, 9]])   # columns come back in DB order
    assert d.load_obs(d.stn_ids[4]).shape == (365,)
    p = str(tmp_path / "db.npz")
    d.save(p)
    d2 = db.StationSerialDataDb(p, "tmin")
    assert np.array_equal(d2.var, d.var) and np.array_equal(d2.stn_ids, d.stn_ids)
    assert np.array_equal(d2.days[db.YMD], d.days[db.YMD])
    for name in d.stns.dtype.names[3:]:
        assert np.array_equal(d2.stns[name], d.stns[name], equal_nan=True)


def test_synthetic_inputs_shape(small_db):
    s = small_db.stns
    assert np.isnan(s[db.BAD]).mean() > 0.9                        # "good" == isnan(bad)
    assert set(np.unique(s[db.get_optim_varname(1)][np.isfinite(s[db.get_optim_varname(1)])])) <= set(synth.NNGH_SET)
    f = synth.Fields()
    w = synth.make_wrk_chk(f, synth.TILE_ROW0, synth.TILE_COL0, 50, 50)
    assert w.shape == (32, 50, 50) and w[2].all()
    assert np.all(np.diff(w[3][:, 0]) < 0) and np.all(np.diff(w[4][0]) > 0)      # lat descends, lon ascends
    assert abs(w[3][0, 0] - (50.0 - (synth.TILE_ROW0 + 0.5) / 120.0)) < 1e-12
    # no exact distance ties between stations (jittered, no lattice)
    d2 = (s[db.LON][:, None] - s[db.LON][None, :]) ** 2 + (s[db.LAT][:, None] - s[db.LAT][None, :]) ** 2
    assert np.unique(d2[np.triu_indices(s.size, 1)]).size == s.size * (s.size - 1) // 2


def _tiler(mask=None):
    ny, nx = 20, 30
    lats = 45.0 - np.arange(ny) * 0.1
    lons = -110.0 + np.arange(nx) * 0.1
    mask = np.ones((ny, nx), bool) if mask is None else mask
    attrs = [("a%d" % i, np.full((ny, nx), float(i)) + np.arange(nx)[None, :]) for i in range(27)]
    return Tiler(dict(mask=mask, lon=lons, lat=lats), attrs, 10, 10, 5, 5), lats, lons


def test_tiler_chunks_match_reference_order():
    t, lats, lons = _tiler()
    assert t.ntiles == 6 and t.ntile_chks == 24 and t.chk_size_i == 32
    assert t.tile_ids[0] == "h00v00" and t.tile_ids[4] == "h01v01"
    assert t.tile_chks[0] == (0, 0, 0, 0, 0) and t.tile_chks[1] == (0, 0, 0, 0, 5) and t.tile_chks[4] == (1, 0, 10, 0, 0)
    k, w = t.next()
    assert k == 0 and w.shape == (32, 5, 5)
    assert np.array_equal(w[0], np.mgrid[0:5, 0:5][0]) and np.array_equal(w[1], np.mgrid[0:5, 0:5][1])
    assert np.all(w[2] == 1) and np.allclose(w[3][:, 0], lats[:5]) and np.allclose(w[4][0], lons[:5])
    assert np.allclose(w[5 + 3], 3.0 + np.arange(5)[None, :])
    n = 1
    for _ in t:
        n += 1
    assert n == 24
    info = t.build_tile_grid_info()
    assert info.chks_per_tile == 4 and info.nchks == 24 and info.get_tile_id(5) == "h02v01"


def test_tiler_skips_empty_tiles_and_partitions():
    mask = np.ones((20, 30), bool)
    mask[:10, 10:20] = False                                        # tile 1 fully masked -> skipped
    mask[10:, :10] = False
    mask[10:15, 0:3] = True                                         # tile with only 15 cells
    t, _, _ = _tiler(mask)
    assert t.ntiles == 5 and sorted(set(c[0] for c in t.tile_chks)) == [0, 1, 2, 3, 4]
    parts = partition_chunks(t.tile_chks, t.mask, 10, 10, 2)
    assert sorted(parts[0] + parts[1]) == sorted(t.tile_chks)
    assert not set(c[0] for c in parts[0]) & set(c[0] for c in parts[1])   # whole tiles
    load = [sum(int(t.mask[c[1] + c[3]:c[1] + c[3] + 5, c[2] + c[4]:c[2] + c[4] + 5].sum()) for c in p) for p in parts]
    assert abs(load[0] - load[1]) <= 100
    assert partition_chunks(t.tile_chks, t.mask, 10, 10, 2, rank=1) == parts[1]
    one = partition_chunks(t.tile_chks, t.mask, 10, 10, 1)
    assert one[0] == t.tile_chks


def test_tiler_rejects_bad_sizes():
    with pytest.raises(ValueError):
        Tiler(dict(mask=np.ones((20, 30), bool), lon=np.arange(30.), lat=np.arange(20.)), [], 7, 10, 5, 5)


def test_oracle_xval_tair_anom_is_gwr_mth_statistics(small_db):
    """oracle.XvalTairAnom (optimize.py:505-545) is bias / MAE / r2 of GwrTairAnom.gwr_mth minus the held-out station's
    own anomalies; check it against a direct evaluation for one station and one neighbour count."""
    from oracle import twx_oracle as o
    from topowx_b200 import db
    oda = o.StationDb(small_db.stns, small_db.var, small_db.days)
    xv = o.XvalTairAnom(oda)
    good = np.nonzero(np.isnan(small_db.stns[db.BAD]) & np.isfinite(small_db.stns[db.MASK]))[0]
    sid = small_db.stns[db.STN_ID][good[7]]
    bias, mae, r2 = xv.run_xval(sid, np.array([40]))
    stn = oda.stns[oda.stn_idxs[sid]]
    for mth in (1, 7):
        obs = oda.load_obs(np.array([sid]))[:, 0].astype(np.float64)[oda.mth_idx[mth]]
        pred = xv.gwr.gwr_mth(stn, mth, 40, stns_rm=sid)
        assert abs(bias[0, mth - 1] - np.mean(pred - obs)) < 1e-12
        assert abs(mae[0, mth - 1] - np.mean(np.abs(pred - obs))) < 1e-12
        assert 0.0 <= r2[0, mth - 1] <= 1.0


def test_build_nstn_bandwidths_matches_reference_set():
    """optimize.py:376-405 evaluated at the arguments of steps 21 and 23."""
    from topowx_b200.interp import build_nstn_bandwidths
    assert build_nstn_bandwidths(35, 150, 0.10).tolist() == [35, 39, 43, 47, 52, 57, 63, 69, 76, 84, 92, 101, 111, 122,
                                                             134, 147]


def test_bench_reference_arm_prints_one_json_line():
    """bench.py --impl reference (the oracle port on the host cores): exactly one line on stdout, carrying the keys of the
    bench contract; anything a library prints on fd 1 must not end up in front of it."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       env=dict(os.environ, TWX_REF_CELLS_PER_STEP="4"), cwd=root, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-1000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "cell-days/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["value"] > 0


def test_tile_writer_roundtrip(tmp_path):
    """TileWriter (tiling.py:304-537): variables, dtypes, fill values, scale factor; chunk-wise and whole-tile writes agree."""
    from scipy.io import netcdf_file
    from topowx_b200 import synth
    from topowx_b200.interp.tiling import Tiler, TileWriter, AsyncTileWriter
    ny, nx = 20, 30
    mask = np.ones((ny, nx), dtype=bool)
    lats, lons = synth.grid_lats(np.arange(ny)), synth.grid_lons(np.arange(nx))
    t = Tiler(dict(mask=mask, lon=lons, lat=lats), [], 10, 10, 5, 5)
    info = t.build_tile_grid_info()
    days = synth.make_days(1995, 1)[:40]
    rng = np.random.default_rng(3)
    tile = dict(tmin=rng.integers(-3000, 3000, (40, 10, 10)).astype(np.int16), tmax=rng.integers(-3000, 3000, (40, 10, 10)).astype(np.int16),
                tmin_norm=rng.normal(size=(12, 10, 10)).astype(np.float32), tmax_norm=rng.normal(size=(12, 10, 10)).astype(np.float32),
                tmin_se=rng.uniform(size=(12, 10, 10)).astype(np.float32), tmax_se=rng.uniform(size=(12, 10, 10)).astype(np.float32),
                ninvalid=rng.integers(0, 5, (10, 10)).astype(np.int32), status=np.zeros((10, 10), np.uint8))
    tw = TileWriter(info, str(tmp_path / "a"))
    os.makedirs(str(tmp_path / "a"))
    tid = info.get_tile_id(1)
    # chunk by chunk, like the writer rank of step25 (step25:229-249); one chunk is left unwritten -> fill values
    for y in range(0, 10, 5):
        for x in range(0, 10, 5):
            if (y, x) == (5, 5):
                continue
            tw.write_tile_chunk(tid, "tmin", days, y, x, tile["tmin"][:, y:y + 5, x:x + 5], tile["tmin_norm"][:, y:y + 5, x:x + 5],
                                tile["tmin_se"][:, y:y + 5, x:x + 5], tile["ninvalid"][y:y + 5, x:x + 5])
    if tw.format.startswith("netCDF3"):
        ds = netcdf_file(os.path.join(str(tmp_path / "a"), tid, "%s_tmin.nc" % tid), "r", mmap=False, maskandscale=False)
        v = ds.variables["tmin"]
        assert v.data.dtype.kind == "i" and v.data.dtype.itemsize == 2 and v.data.shape == (40, 10, 10)
        assert np.array_equal(v.data[:, :5, :], tile["tmin"][:, :5, :]) and (v.data[:, 5:, 5:] == -32767).all()
        assert abs(float(v.scale_factor) - 0.01) < 1e-9 and int(v._FillValue) == -32767
        assert np.array_equal(ds.variables["tmin_normal"].data[:, :5, :5], tile["tmin_norm"][:, :5, :5])
        assert (ds.variables["inconsist_tair"].data[5:, 5:] == -2147483647).all()
        assert ds.variables["time"].data.shape == (40,) and ds.variables["time"].data[0] == 0.5
        r0, c0 = info.tile_rc[tid]
        assert np.allclose(ds.variables["lat"].data, lats[r0:r0 + 10]) and np.allclose(ds.variables["lon"].data, lons[c0:c0 + 10])
        assert ds.variables["climatology_bounds"].data.shape == (12, 2)
        ds.close()
    # whole tiles through the background writer, netCDF and raw
    for fmt in ("nc", "raw"):
        aw = AsyncTileWriter(info, str(tmp_path / fmt), days, fmt=fmt)
        aw.submit(tid, tile)
        nbytes = aw.wait()
        aw.close()
        assert nbytes > tile["tmin"].nbytes * 2
    raw = np.load(os.path.join(str(tmp_path / "raw"), tid, "%s_tmax.npy" % tid))
    assert np.array_equal(raw, tile["tmax"])


def test_predictor_store_and_tile_feed(tmp_path):
    """rasters on disk -> Tiler -> TileFeed (background reader, ring of buffers): every tile arrives once,
    in the reference's order, with the planes Tiler.next would build."""
    from topowx_b200 import synth
    from topowx_b200.interp import Tiler, PredictorStore, TileFeed
    f = synth.Fields()
    store = PredictorStore.create_synthetic(str(tmp_path / "rasters"), f, 240, 1230, 40, 60, band=16)
    tiler = Tiler(store, store.attrs(), 20, 20, 20, 20)
    assert tiler.chk_size_i == 32
    ref = synth.make_wrk_chk_grid(f, 240, 1230, 40, 60)
    seen = []
    feed = TileFeed(tiler, depth=3, pinned=False)
    for k, tid, wrk in feed:
        _, i, j, y, x = tiler.tile_chks[len(seen)]
        assert tid == tiler.tile_ids[k]
        w = wrk.numpy()
        np.testing.assert_array_equal(w[2], ref[2, i:i + 20, j:j + 20])
        np.testing.assert_allclose(w[3:], ref[3:, i:i + 20, j:j + 20], rtol=0, atol=1e-9)
        assert w[0, 3, 5] == 3 and w[1, 3, 5] == 5
        seen.append(k)
    assert seen == [c[0] for c in tiler.tile_chks] and len(seen) == tiler.ntiles
    assert feed.bytes_read == len(seen) * 32 * 400 * 8


def test_station_data_wrk_chk_cache():
    """StationDataWrkChk.set_obs / load_obs (interp_tair.py:1027-1097): DB-order columns, buffer grows until all found."""
    from topowx_b200 import synth
    from topowx_b200.interp import StationDataWrkChk
    f = synth.Fields()
    days = synth.make_days(1995, 1)
    da = synth.make_station_db(0, 300, synth.tile_bbox(buf=4.0), f, days)
    wc = StationDataWrkChk((da.stns, da.var, da.days), "tmin")
    b = synth.tile_bbox(buf=0.0)
    wc.set_obs(b, deg_buf=1)
    n1 = wc.chk_stnids.size
    assert 0 < n1 < 300
    inside = wc.chk_stnids[[5, 1, 9]]
    got = wc.load_obs(inside, mth=3)
    want = da.load_obs(np.sort(inside), mth=3)
    np.testing.assert_array_equal(got, want)                      # DB (= id) order, whatever the request order
    far = np.setdiff1d(da.stn_ids, wc.chk_stnids)[:2]
    got = wc.load_obs(np.concatenate([inside, far]), mth=7)
    assert wc.chk_deg_buf > 1 and got.shape == (31, 5)
    np.testing.assert_array_equal(got, da.load_obs(np.sort(np.concatenate([inside, far])), mth=7))


def test_predictor_grids_lookup(tmp_path):
    """PredictorGrids.setPtValues (interp_tair.py:87-141) over a PredictorStore: order 0 = value of the containing cell and the
    point moved to the cell centre; order 1 = bilinear."""
    from topowx_b200 import synth, db
    from topowx_b200.interp import PredictorStore, PredictorGrids
    from topowx_b200.interp.interp_tair import build_empty_pt
    f = synth.Fields()
    store = PredictorStore.create_synthetic(str(tmp_path / "r"), f, 1000, 3000, 30, 40)
    pg = PredictorGrids(store, interpOrders={"elev": 1})
    ref = synth.make_wrk_chk_grid(f, 1000, 3000, 30, 40)
    pt = build_empty_pt()
    lat_c, lon_c = synth.grid_lats(1012), synth.grid_lons(3025)
    pt[db.LON], pt[db.LAT] = lon_c + 0.003, lat_c - 0.002             # inside cell (12, 25)
    pg.setPtValues(pt, chgLatLon=True)
    assert pt[db.LON] == lon_c and pt[db.LAT] == lat_c
    assert pt[db.ELEV] == ref[5, 12, 25] and pt[db.TDI] == ref[6, 12, 25] and pt[db.CLIMDIV] == ref[7, 12, 25]
    assert pt["tmin03"] == ref[8 + 2, 12, 25] and pt["tmax11"] == ref[20 + 10, 12, 25] and pt[db.MASK] == ref[2, 12, 25]
    pt[db.LON], pt[db.LAT] = lon_c + 0.5 * synth.RES, lat_c           # half way to the next column, order 1
    pg.setPtValues(pt, chgLatLon=False)
    assert abs(pt[db.ELEV] - 0.5 * (ref[5, 12, 25] + ref[5, 12, 26])) < 1e-9
    assert pt[db.TDI] in (ref[6, 12, 25], ref[6, 12, 26])
    import pytest
    pt[db.LON] = lon_c + 10.0
    with pytest.raises(Exception):
        pg.setPtValues(pt)


def test_ked_row_offset_table_matches_the_tile_indexing():
    """csrc/ked.cu looks the byte offset of tile row I up in c_ked_rowoff instead of computing ltile(I, 0) * 512; the table
    must cover every row the largest size class can touch (rows 0 .. KED_NBMAX, the last one being the augmented B' row)."""
    import os, re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "topowx_b200", "csrc", "ked.cu")).read()
    m = re.search(r"c_ked_rowoff\[(\d+)\]\s*=\s*\{([^}]*)\}", src)
    vals = [int(v) for v in m.group(2).split(",")]
    assert len(vals) == int(m.group(1))
    assert vals == [i * (i - 1) // 2 * 512 for i in range(len(vals))]
    hdr = open(os.path.join(root, "include", "twxi.h")).read()
    max_n = int(re.search(r"#define\s+TWXI_MAX_KRIG_NNGHS\s+(\d+)", hdr).group(1))
    assert len(vals) > max_n // 8 + 1
