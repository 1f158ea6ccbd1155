"""The twx.interp-compatible Python API on the GPU path against the oracle's classes of the same names, plus
size-independent properties at the benchmark's full tile size.  Run on the B200 box with -m gpu."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import twx_oracle as o          # noqa: E402  (checker only)

TOL_C = 1e-4


@pytest.fixture(scope="module")
def env():
    from topowx_b200 import synth, db
    f = synth.Fields()
    days = synth.make_days(1995, 1)
    da = [synth.make_station_db(w, 2000, synth.tile_bbox(), f, days) for w in (0, 1)]
    oda = [o.StationDb(d.stns, d.var, d.days) for d in da]
    return dict(f=f, da=da, oda=oda, synth=synth, db=db, days=days)


def _fill_pt(env, pt, row, col, names):
    synth, f = env["synth"], env["f"]
    lat, lon = float(synth.grid_lats(row)), float(synth.grid_lons(col))
    elev = float(f.elev(lon, lat))
    pt[names.LAT], pt[names.LON], pt[names.ELEV], pt[names.TDI] = lat, lon, elev, float(f.tdi(lon, lat))
    pt[names.CLIMDIV] = float(f.climdiv(lon, lat))
    for m in range(1, 13):
        pt["tmin%02d" % m] = float(f.lst(0, m, lon, lat, elev))
        pt["tmax%02d" % m] = float(f.lst(1, m, lon, lat, elev))
    return pt


def test_station_select_api(env):
    from topowx_b200.interp import StationSelect
    db = env["db"]
    da, oda = env["da"][0], env["oda"][0]
    good = np.isnan(da.stns[db.BAD])
    ss, oss = StationSelect(da, good), o.StationSelect(oda, good)
    for lat, lon, nn, mth in [(40.61, -99.02, 35, 1), (39.9, -98.3, 100, 7), (41.2, -99.7, 147, None)]:
        ss.set_ngh_stns(lat, lon, nn, load_obs=True, obs_mth=mth)
        oss.set_ngh_stns(lat, lon, nn, load_obs=True, obs_mth=mth)
        assert np.array_equal(ss.ngh_stns[db.STN_ID], oss.ngh_stns[db.STN_ID])     # same stations, id order
        np.testing.assert_allclose(ss.ngh_dists, oss.ngh_dists, rtol=1e-12)
        np.testing.assert_allclose(ss.ngh_wgt, oss.ngh_wgt, rtol=0, atol=1e-11)
        assert np.array_equal(ss.ngh_obs, oss.ngh_obs)
    sid = da.stn_ids[good][17]
    st = da.stns[da.stn_idxs[sid]]
    ss0 = StationSelect(da, good, rm_zero_dist_stns=True)
    ss0.set_ngh_stns(st[db.LAT], st[db.LON], 50, load_obs=False, stns_rm=sid)
    assert sid not in ss0.ngh_stns[db.STN_ID] and ss0.ngh_obs is None
    # more neighbours than candidate stations -> IndexError like station_select.py:164
    few = np.zeros(da.stns.size, dtype=bool)
    few[np.nonzero(good)[0][:30]] = True
    ss_few = StationSelect(da, few)
    with pytest.raises(IndexError):
        ss_few.set_ngh_stns(40.0, -99.0, 30, load_obs=False)
    ss_few.set_ngh_stns(40.0, -99.0, 29, load_obs=False)
    assert ss_few.ngh_stns.size == 29


def test_krig_gwr_interp_classes(env):
    from topowx_b200.interp import StationSelect, KrigTair, GwrTairAnom, InterpTair
    from topowx_b200.interp.interp_tair import build_empty_pt
    db = env["db"]
    da, oda = env["da"][1], env["oda"][1]
    good = np.isnan(da.stns[db.BAD])
    ss = StationSelect(da, good)
    kt, gt = KrigTair(ss), GwrTairAnom(ss)
    oss = o.StationSelect(oda, good)
    okt, ogt = o.KrigTair(oss), o.GwrTairAnom(oss)
    pt, opt = _fill_pt(env, build_empty_pt(), 1100, 3111, db), o.build_empty_pt()
    for n in opt.dtype.names:
        opt[n] = pt[n] if n in pt.dtype.names else 0.0
    for m in (1, 8):
        pt[db.get_lst_varname(m)] = pt["tmax%02d" % m]
        opt[o.lst_name(m)] = pt["tmax%02d" % m]
        mean, var = kt.krig(pt, m)
        omean, ovar = okt.krig(opt, m)
        assert abs(mean - omean) < TOL_C and abs(var - ovar) <= 1e-6 * abs(ovar)
        se, ci = kt.std_err_ci(mean, var)
        ose, oci = okt.std_err_ci(omean, ovar)
        assert abs(se - ose) < 1e-7 and abs(ci[0] - oci[0]) < TOL_C
        pt[db.get_norm_varname(m)] = mean
        opt[o.norm_name(m)] = omean
        assert np.abs(gt.gwr_mth(pt, m) - ogt.gwr_mth(opt, m)).max() < TOL_C
        assert np.abs(gt.gwr_mth(pt, m, nnghs=40) - ogt.gwr_mth(opt, m, nnghs=40)).max() < TOL_C
    it, oit = InterpTair(kt, gt), o.InterpTair(okt, ogt)
    for m in range(1, 13):
        pt[db.get_lst_varname(m)] = pt["tmax%02d" % m]
        opt[o.lst_name(m)] = pt["tmax%02d" % m]
    dly, norms, se = it.interp(pt)
    odly, onorms, ose = oit.interp(opt)
    assert np.abs(dly - odly).max() < TOL_C and np.abs(norms - onorms).max() < TOL_C and np.abs(se - ose).max() < 1e-6
    assert abs(pt[db.get_norm_varname(5)] - norms[4]) == 0          # side effect kept (interp_tair.py:433)


def test_pt_interp_and_xval_classes(env):
    from topowx_b200.interp import PtInterpTair, XvalTairOverall
    db = env["db"]
    pti = PtInterpTair(env["da"][0], env["da"][1])
    opti = o.PtInterpTair(env["oda"][0], env["oda"][1])
    assert pti.days.size == 365
    _fill_pt(env, pti.a_pt, 1203, 3040, db)
    for n in opti.a_pt.dtype.names:
        opti.a_pt[n] = pti.a_pt[n]
    r, orr = pti.interp_pt(), opti.interp_pt()
    for a, b in zip(r[:6], orr[:6]):
        assert np.abs(np.asarray(a) - np.asarray(b)).max() < TOL_C
    assert r[6] == orr[6]
    pti.a_pt[db.CLIMDIV] = 12345.0
    with pytest.raises(KeyError):
        pti.interp_pt()
    xv, oxv = XvalTairOverall(env["da"][0], "tmin"), o.XvalTairOverall(env["oda"][0])
    da = env["da"][0]
    cand = da.stn_ids[np.isnan(da.stns[db.BAD]) & np.isfinite(da.stns[db.MASK])]
    sid = cand[123]
    d, n, s = xv.run_interp(sid)
    od, on, os_ = oxv.run_interp(sid)
    assert np.abs(d - od).max() < TOL_C and np.abs(n - on).max() < TOL_C and np.abs(s - os_).max() < 1e-6
    dly, norms, se, st = xv.run_interp_batch(cand[120:126])
    assert np.all(st == 0) and np.abs(dly[3] - od).max() < TOL_C and np.abs(norms[3] - on).max() < TOL_C


def test_full_tile_properties_and_sampled_parity(env):
    """BASELINE config 2 at full size: 250x250 tile x 365 days.  Properties: every cell succeeds, splitting the
    tile into 50x50 work chunks gives byte-identical output (partition invariance = multi-GPU invariance),
    Tmin < Tmax everywhere after the fixer, quantised values round-trip the normals; plus a random sample of
    cells against the oracle."""
    from topowx_b200.context import TwxiContext, interp_chunk
    synth, db = env["synth"], env["db"]
    ctx = [TwxiContext(d, np.isnan(d.stns[db.BAD])) for d in env["da"]]
    wrk = synth.make_wrk_chk(env["f"], synth.TILE_ROW0, synth.TILE_COL0, 250, 250)
    out = interp_chunk(ctx[0], ctx[1], wrk)
    assert np.all(out["status"] == 0)
    assert np.all(out["tmin"] <= out["tmax"])                      # equal only after 0.01 C quantisation
    assert out["tmin"].min() > -6000 and out["tmax"].max() < 6000
    assert np.all(out["tmin_se"] > 0) and np.all(out["tmax_se"] < 5)
    # partition invariance on two chunks
    for (y, x) in [(0, 0), (150, 200)]:
        sub = np.ascontiguousarray(wrk[:, y:y + 50, x:x + 50])
        o2 = interp_chunk(ctx[0], ctx[1], sub)
        for k in ("tmin", "tmax", "tmin_norm", "tmax_norm", "tmin_se", "tmax_se", "ninvalid", "status"):
            assert np.array_equal(o2[k], out[k][..., y:y + 50, x:x + 50]), k
    # monthly means of the daily output track the kriged normals (GWR anomalies average to ~0 over a month
    # only approximately; this is a sanity bound, not parity)
    mth = np.asarray(env["days"][db.MONTH])
    mm = np.stack([out["tmax"][mth == m].mean(axis=0) * 0.01 for m in range(1, 13)])
    assert np.abs(mm - out["tmax_norm"]).max() < 12.0
    # sampled parity against the oracle
    pti = o.PtInterpTair(env["oda"][0], env["oda"][1])
    r = np.random.default_rng(11)
    cells = [(int(a), int(b)) for a, b in zip(r.integers(0, 250, 12), r.integers(0, 250, 12))]
    ref = o.interp_chunk(pti, wrk, cells=cells)
    nfix = 0
    for (a, b) in cells:
        assert ref["status"][a, b] == 0
        for k in ("tmin", "tmax"):
            d = np.abs(out[k][:, a, b].astype(int) - ref[k][:, a, b].astype(int))
            assert d.max() <= 1 and (d > 0).sum() <= 2
        for k in ("tmin_norm", "tmax_norm", "tmin_se", "tmax_se"):
            assert np.abs(out[k][:, a, b] - ref[k][:, a, b]).max() < TOL_C
        assert out["ninvalid"][a, b] == ref["ninvalid"][a, b]
        nfix += int(ref["ninvalid"][a, b])
    print("cells with Tmin>=Tmax fixes in the tile:", int((out["ninvalid"] > 0).sum()), "sampled fixes:", nfix)


def test_fixer_with_many_inversions():
    """Tmin >= Tmax on many days (synthetic diurnal range forced to 0.3 C): the GPU fixer, the 1981-2010
    normals recomputation and ninvalid against the oracle (interp_tair.py:143-197, 583-590)."""
    from topowx_b200 import synth, db
    from topowx_b200.context import TwxiContext, interp_chunk, interp_cells
    f = synth.Fields(dtr_override=0.3)
    days = synth.make_days(1995, 1)
    da = [synth.make_station_db(w, 700, synth.tile_bbox(buf=2.0), f, days, seed=77) for w in (0, 1)]
    # same station set for both variables so that the interpolated fields really cross
    ctx = [TwxiContext(d, np.isnan(d.stns[db.BAD])) for d in da]
    wrk = synth.make_wrk_chk(f, synth.TILE_ROW0 + 10, synth.TILE_COL0 + 20, 3, 4)
    out = interp_chunk(ctx[0], ctx[1], wrk)
    pti = o.PtInterpTair(*[o.StationDb(d.stns, d.var, d.days) for d in da])
    ref = o.interp_chunk(pti, wrk)
    assert np.array_equal(out["status"], ref["status"]) and np.all(ref["status"] == 0)
    assert ref["ninvalid"].min() > 5                                   # the fixer really ran
    assert np.array_equal(out["ninvalid"], ref["ninvalid"])
    for k in ("tmin", "tmax"):
        d = np.abs(out[k].astype(int) - ref[k].astype(int))
        assert d.max() <= 1 and (d > 0).mean() < 2e-3
    for k in ("tmin_norm", "tmax_norm"):                               # recomputed from the fixed dailies
        assert np.abs(out[k] - ref[k]).max() < TOL_C
    assert np.all(out["tmin"] <= out["tmax"])                      # equal only after 0.01 C quantisation
    # float64 cells API, with and without the fixer
    lat, lon = wrk[3].ravel(), wrk[4].ravel()
    lst_a = np.stack([wrk[8 + m].ravel() for m in range(12)], axis=1)
    lst_b = np.stack([wrk[20 + m].ravel() for m in range(12)], axis=1)
    r = interp_cells(ctx[0], ctx[1], lat, lon, wrk[5].ravel(), wrk[6].ravel(), wrk[7].ravel(), lst_a, lst_b)
    for i, (a, b) in enumerate([(a, b) for a in range(3) for b in range(4)]):
        assert np.abs(r[0][i] - ref["tmin_f8"][(a, b)]).max() < TOL_C
        assert np.abs(r[1][i] - ref["tmax_f8"][(a, b)]).max() < TOL_C
    r0 = interp_cells(ctx[0], ctx[1], lat, lon, wrk[5].ravel(), wrk[6].ravel(), None, lst_a, lst_b, fix_invalid=False)
    assert (r0[0] >= r0[1]).sum() == ref["ninvalid"].sum() and np.all(r0[6] == 0)


def test_fixer_window_without_valid_day():
    """Tmax mostly below Tmin: windows with no valid day -> 'No valid tmin/tmax in window' (status 5), the cell
    keeps fill values (interp_tair.py:191-192, step25:154-160)."""
    from topowx_b200 import synth, db
    from topowx_b200.context import TwxiContext, interp_chunk
    f = synth.Fields(dtr_override=-1.2)
    days = synth.make_days(1995, 1)
    da = [synth.make_station_db(w, 700, synth.tile_bbox(buf=2.0), f, days, seed=78) for w in (0, 1)]
    ctx = [TwxiContext(d, np.isnan(d.stns[db.BAD])) for d in da]
    wrk = synth.make_wrk_chk(f, synth.TILE_ROW0 + 30, synth.TILE_COL0 + 40, 2, 3)
    out = interp_chunk(ctx[0], ctx[1], wrk)
    pti = o.PtInterpTair(*[o.StationDb(d.stns, d.var, d.days) for d in da])
    ref = o.interp_chunk(pti, wrk)
    assert np.array_equal(out["status"], ref["status"])
    assert (ref["status"] == o.ST_FIXER_EMPTY).any()
    bad = ref["status"] != 0
    assert np.all(out["tmin"][:, bad] == -32767) and np.all(out["ninvalid"][bad] == -2147483647)
    assert np.array_equal(out["ninvalid"], ref["ninvalid"])


def test_grid_knn_candidate_pruning_is_exact(env, monkeypatch):
    """Work chunks search neighbours over per-block candidate lists (csrc/knn.cu, knn_candidates_kernel).  The lists
    are supersets of every cell's k+1 nearest stations, so the chunk output must be byte-identical to the full-table
    scan (TWXI_KNN_FULL=1) — on a ragged chunk (sides not multiples of the 25-cell block) with masked cells."""
    from topowx_b200.context import TwxiContext, interp_chunk
    synth, db = env["synth"], env["db"]
    ctx = [TwxiContext(d, np.isnan(d.stns[db.BAD])) for d in env["da"]]
    wrk = synth.make_wrk_chk(env["f"], synth.TILE_ROW0 + 20, synth.TILE_COL0 + 40, 60, 83)
    wrk[2, 5:9, :] = 0                                              # masked rows
    wrk[2, :, 30] = 0
    pruned = interp_chunk(ctx[0], ctx[1], wrk)
    monkeypatch.setenv("TWXI_KNN_FULL", "1")
    full = interp_chunk(ctx[0], ctx[1], wrk)
    monkeypatch.delenv("TWXI_KNN_FULL")
    assert (pruned["status"] == 255).sum() == 4 * 83 + 60 - 4
    for k in ("tmin", "tmax", "tmin_norm", "tmax_norm", "tmin_se", "tmax_se", "ninvalid", "status"):
        assert np.array_equal(pruned[k], full[k]), k
    # candidate lists that overflow their capacity: those blocks are searched over the whole table by the second launch
    # (knn_kernel mode 1); 230 entries overflow some blocks of this chunk and not others, 64 overflows all of them
    for cap in ("230", "64"):
        monkeypatch.setenv("TWXI_KNN_CAP", cap)
        capped = interp_chunk(ctx[0], ctx[1], wrk)
        monkeypatch.delenv("TWXI_KNN_CAP")
        for k in ("tmin", "tmax", "tmin_norm", "tmax_norm", "tmin_se", "tmax_se", "ninvalid", "status"):
            assert np.array_equal(capped[k], full[k]), (cap, k)


def test_xval_tair_anom_class(env):
    """XvalTairAnom (optimize.py:477-545, the step23 driver): leave-one-out GWR at a station for a set of neighbour
    counts -> bias / MAE / r2 per (count, month); GPU mirror (single and batch form) against the oracle."""
    from topowx_b200 import interp as twx_interp
    db = env["db"]
    w = 0
    da, oda = env["da"][w], env["oda"][w]
    xv_gpu = twx_interp.XvalTairAnom(da, "tmin")
    xv_ref = o.XvalTairAnom(oda)
    cand = np.nonzero(np.isnan(da.stns[db.BAD]) & np.isfinite(da.stns[db.MASK]))[0]
    ids = da.stns[db.STN_ID][cand[np.random.default_rng(21).choice(cand.size, 3, replace=False)]]
    a_nnghs = np.array([35, 63, 147])
    bias, mae, r2, st = xv_gpu.run_xval_batch(ids, a_nnghs)
    assert np.all(st == 0)
    for i, sid in enumerate(ids):
        rb, rm_, rr = xv_ref.run_xval(sid, a_nnghs)
        assert np.abs(bias[i] - rb).max() < 1e-6 and np.abs(mae[i] - rm_).max() < 1e-6
        assert np.abs(r2[i] - rr).max() < 1e-8
    b1, m1, r1 = xv_gpu.run_xval(ids[0], a_nnghs)                  # reference signature
    assert np.array_equal(b1, bias[0]) and np.array_equal(m1, mae[0]) and np.array_equal(r1, r2[0])


def test_async_chunks_equal_synchronous_calls(env):
    """twxi_interp_chunk_async: three chunks submitted back to back from pinned host buffers (staging slots are reused
    from the third on), results byte-identical to the synchronous call."""
    import torch
    from topowx_b200.context import TwxiContext, interp_chunk, interp_chunk_wait
    synth, f, db = env["synth"], env["f"], env["db"]
    ctx = [TwxiContext(d, np.isnan(d.stns[db.BAD])) for d in env["da"]]
    chunks = [synth.make_wrk_chk(f, synth.TILE_ROW0 + r, synth.TILE_COL0 + c, 6, 7) for r, c in ((3, 5), (120, 40), (200, 210))]
    chunks[1][2, 2:4, 1:3] = 0                                   # some masked cells
    ref = [interp_chunk(ctx[0], ctx[1], w) for w in chunks]
    nd = ctx[0].ndays

    def pinned():
        mk = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
        return dict(tmin=mk((nd, 6, 7), torch.int16), tmax=mk((nd, 6, 7), torch.int16),
                    tmin_norm=mk((12, 6, 7), torch.float32), tmax_norm=mk((12, 6, 7), torch.float32),
                    tmin_se=mk((12, 6, 7), torch.float32), tmax_se=mk((12, 6, 7), torch.float32),
                    ninvalid=mk((6, 7), torch.int32), status=mk((6, 7), torch.uint8))
    outs = [pinned() for _ in chunks]
    wrks = [torch.from_numpy(w).pin_memory() for w in chunks]
    for w, out in zip(wrks, outs):
        interp_chunk(ctx[0], ctx[1], w, out=out, wait=False)
    interp_chunk_wait(ctx[0])
    for r, out in zip(ref, outs):
        for k in r:
            np.testing.assert_array_equal(np.asarray(r[k]), out[k].numpy(), err_msg=k)
    # and the synchronous call still works afterwards
    again = interp_chunk(ctx[0], ctx[1], chunks[0])
    for k in again:
        np.testing.assert_array_equal(np.asarray(ref[0][k]), np.asarray(again[k]), err_msg=k)
