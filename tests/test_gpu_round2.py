"""Round-2 GPU tests: wider parity samples, the cases the round-1 review found untested, the async contract, contexts on
two streams, leave-outs through PtInterpTair.interp_pt, per-point limits.  Run on the B200 box with -m gpu."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import twx_oracle as o          # noqa: E402  (checker only)
from oracle import cpu_farm                 # noqa: E402

TOL_C = 1e-4


@pytest.fixture(scope="module")
def env():
    from topowx_b200 import synth, db
    f = synth.Fields()
    days = synth.make_days(1995, 1)
    da = [synth.make_station_db(w, 2000, synth.tile_bbox(), f, days) for w in (0, 1)]
    oda = [o.StationDb(d.stns, d.var, d.days) for d in da]
    return dict(f=f, da=da, oda=oda, synth=synth, db=db, days=days)


def _check_cells(out, ref, cells):
    nfix = 0
    for rc in cells:
        st, ninv, vals = ref[rc]
        assert out["status"][rc] == st, rc
        if st != 0:
            continue
        for k in ("tmin", "tmax"):
            d = np.abs(out[k][:, rc[0], rc[1]].astype(int) - vals[k].astype(int))
            assert d.max() <= 1 and (d > 0).sum() <= 3, (rc, k)          # +-1 count only at x.xx5 boundaries
        for k in ("tmin_norm", "tmax_norm", "tmin_se", "tmax_se"):
            assert np.abs(out[k][:, rc[0], rc[1]] - vals[k]).max() < TOL_C, (rc, k)
        assert out["ninvalid"][rc] == ninv, rc
        nfix += ninv
    return nfix


def test_full_tile_256_cells_vs_oracle(env):
    """configs[1] at full size against the oracle on 256 random cells (the oracle runs on all host cores)."""
    from topowx_b200.context import TwxiContext, interp_chunk
    synth, db = env["synth"], env["db"]
    ctx = [TwxiContext(d, np.isnan(d.stns[db.BAD])) for d in env["da"]]
    wrk = synth.make_wrk_chk(env["f"], synth.TILE_ROW0, synth.TILE_COL0, 250, 250)
    out = interp_chunk(ctx[0], ctx[1], wrk)
    r = np.random.default_rng(2025)
    flat = r.choice(250 * 250, 256, replace=False)
    cells = [(int(i // 250), int(i % 250)) for i in flat]
    ref = cpu_farm.interp_cells_parallel(env["da"][0], env["da"][1], wrk, cells)
    assert len(ref) == 256
    _check_cells(out, ref, cells)


def test_conus_scale_64_cells_and_32_loo_stations():
    """C3-C5 station table (10 000 stations over CONUS, land mask, out-of-domain stations): 64 cells of a coastal chunk
    (masked cells, NaN optim / variogram columns nearby) and 32 leave-one-out stations against the oracle."""
    from topowx_b200 import synth, db
    from topowx_b200.context import TwxiContext, interp_chunk
    f = synth.Fields()
    days = synth.make_days(1995, 1)
    da = [synth.make_station_db(w, 10000, synth.conus_bbox(), f, days) for w in (0, 1)]
    oda = [o.StationDb(d.stns, d.var, d.days) for d in da]
    ctx = [TwxiContext(d, np.isnan(d.stns[db.BAD])) for d in da]
    wrk = synth.make_wrk_chk_grid(f, 500, 1250, 250, 250)          # tile with 49 002 land cells: coast inside
    assert 0 < (wrk[2] != 0).sum() < 62500
    out = interp_chunk(ctx[0], ctx[1], wrk)
    assert np.all(out["status"][wrk[2] == 0] == 255) and np.all(out["tmin"][:, wrk[2] == 0] == -32767)
    land = np.argwhere(wrk[2] != 0)
    r = np.random.default_rng(5)
    cells = [tuple(int(v) for v in land[i]) for i in r.choice(len(land), 64, replace=False)]
    ref = cpu_farm.interp_cells_parallel(da[0], da[1], wrk, cells)
    _check_cells(out, ref, cells)
    w = 1
    xv = o.XvalTairOverall(oda[w])
    cand = np.nonzero(np.isnan(da[w].stns[db.BAD]) & np.isfinite(da[w].stns[db.MASK]))[0]
    sel = cand[r.choice(cand.size, 32, replace=False)]
    s = da[w].stns[sel]
    lst = np.stack([s[db.get_lst_varname(m)] for m in range(1, 13)], axis=1)
    rm = ctx[w].local_of_db[sel].astype(np.int32).reshape(-1, 1)
    dly, norms, se, var, st = ctx[w].interp_points(s[db.LAT], s[db.LON], s[db.ELEV], s[db.TDI], lst, rm_idx=rm, rm_zero=True,
                                                   daily=False)
    for i, sid in enumerate(s[db.STN_ID]):
        try:
            od, on, ose = xv.run_interp(sid, norms_only=True)
        except o.OracleError as e:
            assert st[i] == e.status
            continue
        assert st[i] == 0
        assert np.abs(on - norms[i]).max() < TOL_C and np.abs(ose - se[i]).max() < 1e-6


def test_prediction_at_a_station_without_rm_zero(env):
    """A prediction point co-located with a station, rm_zero_dist off, the point's predictors equal to the station's:
    gstat's exact interpolator (nugget at h == 0) returns the station's normal with zero variance."""
    from topowx_b200.context import TwxiContext
    db = env["db"]
    da, oda = env["da"][1], env["oda"][1]
    good = np.isnan(da.stns[db.BAD])
    ctx = TwxiContext(da, good)
    dom = np.nonzero(good & np.isfinite(da.stns[db.MASK]))[0]
    okt = o.KrigTair(o.StationSelect(oda, good))
    for gi in dom[[5, 77, 300]]:
        s = da.stns[gi]
        lst = np.array([s[db.get_lst_varname(m)] for m in range(1, 13)])
        mean, var, st = ctx.krig(s[db.LAT], s[db.LON], s[db.ELEV], lst)
        assert st[0] == 0
        want = np.array([s[db.get_norm_varname(m)] for m in range(1, 13)])
        assert np.abs(mean[0] - want).max() < 1e-8, np.abs(mean[0] - want).max()
        assert np.abs(var[0]).max() < 1e-8
        pt = o.build_empty_pt()
        pt[o.LAT], pt[o.LON], pt[o.ELEV] = s[db.LAT], s[db.LON], s[db.ELEV]
        for m in range(1, 13):
            pt[o.lst_name(m)] = lst[m - 1]
        om, ov = okt.krig(pt, 3)
        assert abs(om - mean[0, 2]) < 1e-8 and abs(ov - var[0, 2]) < 1e-8


def test_chunk_cells_with_too_few_stations():
    """Fewer than 101 candidate stations: set_ngh_stns(100) raises IndexError for every cell (station_select.py:164); the
    chunk call reports TWXI_ST_TOO_FEW_STNS per cell and leaves fill values, the call itself succeeds."""
    from topowx_b200 import synth, db, _lib
    from topowx_b200.context import TwxiContext, interp_chunk
    f = synth.Fields()
    days = synth.make_days(1995, 1)
    da = [synth.make_station_db(w, 90, synth.tile_bbox(buf=1.0), f, days, frac_bad=0.0) for w in (0, 1)]
    ctx = [TwxiContext(d, np.isnan(d.stns[db.BAD])) for d in da]
    wrk = synth.make_wrk_chk(f, synth.TILE_ROW0, synth.TILE_COL0, 4, 5)
    wrk[2, 0, 0] = 0
    out = interp_chunk(ctx[0], ctx[1], wrk)
    assert out["status"][0, 0] == _lib.ST_MASKED
    assert np.all(out["status"].ravel()[1:] == _lib.ST_TOO_FEW_STNS)
    assert np.all(out["tmin"] == _lib.FILL_I2) and np.all(out["tmax_norm"] == _lib.FILL_F4)
    assert np.all(out["ninvalid"] == _lib.FILL_I4)


def test_near_tie_distances_give_the_same_neighbour_set():
    """Station pairs placed symmetrically about the query point: their haversine distances agree to ~1e-13 relative, where
    libdevice and libm may order them differently.  The neighbour SET must still equal the oracle's (the order inside a
    near-tie may differ by rounding of sin)."""
    from topowx_b200 import synth, db
    from topowx_b200.context import TwxiContext
    f = synth.Fields()
    days = synth.make_days(1995, 1)[:31]
    da = synth.make_station_db(0, 400, synth.tile_bbox(buf=1.0), f, days, frac_bad=0.0)
    lat0, lon0 = 40.5, -99.0
    stns = da.stns.copy()
    # mirror the first 150 stations through the query point in longitude: same |dlon|, same latitude -> tie up to rounding
    stns[db.LON][150:300] = 2 * lon0 - stns[db.LON][:150]
    stns[db.LAT][150:300] = stns[db.LAT][:150]
    da2 = db.StationSerialDataDb((stns, da.var, days), "tmin")
    ctx = TwxiContext(da2, None, with_obs=False)
    oss = o.StationSelect(o.StationDb(stns, da.var, days), None)
    for k in (35, 60, 100):
        idx, dist, wgt, st = ctx.knn(lat0, lon0, k)
        assert st[0] == 0
        oss.set_ngh_stns(lat0, lon0, k, load_obs=False)
        od = oss.pt_sort_stn_dists
        # a tie straddling the selection boundary makes the set itself ambiguous: skip those k
        if abs(od[k - 1] - od[k]) <= 1e-9 * od[k]:
            continue
        assert set(idx[0, :k].tolist()) == set(oss.ngh_idx.tolist())
        np.testing.assert_allclose(np.sort(dist[0, :k]), np.sort(oss.ngh_dists), rtol=1e-12)


def test_async_results_valid_after_two_further_submissions(env):
    """The documented contract of twxi_interp_chunk_async: once chunk t+2 has been submitted, chunk t's results are
    complete in the caller's (pinned) buffers - read here WITHOUT the final wait."""
    import torch
    from topowx_b200.context import TwxiContext, interp_chunk, interp_chunk_wait
    synth, f, db = env["synth"], env["f"], env["db"]
    ctx = [TwxiContext(d, np.isnan(d.stns[db.BAD])) for d in env["da"]]
    pos = [(3, 5), (120, 40), (200, 210), (60, 61), (10, 200)]
    chunks = [synth.make_wrk_chk(f, synth.TILE_ROW0 + r, synth.TILE_COL0 + c, 20, 20) for r, c in pos]
    ref = [interp_chunk(ctx[0], ctx[1], w) for w in chunks]
    nd = ctx[0].ndays
    mk = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
    outs = [dict(tmin=mk((nd, 20, 20), torch.int16), tmax=mk((nd, 20, 20), torch.int16),
                 tmin_norm=mk((12, 20, 20), torch.float32), tmax_norm=mk((12, 20, 20), torch.float32),
                 tmin_se=mk((12, 20, 20), torch.float32), tmax_se=mk((12, 20, 20), torch.float32),
                 ninvalid=mk((20, 20), torch.int32), status=mk((20, 20), torch.uint8)) for _ in range(3)]
    wrks = [torch.from_numpy(w).pin_memory() for w in chunks]
    for t in range(len(chunks)):
        if t >= 3:
            for k in outs[t % 3]:
                outs[t % 3][k].fill_(0)                              # the buffer set of chunk t-3 is ours again: reuse it
        interp_chunk(ctx[0], ctx[1], wrks[t], out=outs[t % 3], wait=False)
        if t >= 2:                                                   # chunk t-2 must be complete now
            for k, v in ref[t - 2].items():
                np.testing.assert_array_equal(np.asarray(v), outs[(t - 2) % 3][k].numpy(), err_msg="chunk %d %s" % (t - 2, k))
    interp_chunk_wait(ctx[0])
    for t in (len(chunks) - 2, len(chunks) - 1):
        for k, v in ref[t].items():
            np.testing.assert_array_equal(np.asarray(v), outs[t % 3][k].numpy(), err_msg=k)


def test_two_context_pairs_on_two_streams_from_one_thread(env):
    """Device workspaces belong to the context: two context pairs with different station tables, driven from one host thread
    on two CUDA streams with device-resident buffers (calls return after enqueueing), give the results of running each
    alone."""
    import torch
    from topowx_b200.context import TwxiContext, interp_chunk
    synth, f, db = env["synth"], env["f"], env["db"]
    days = env["days"]
    da_b = [synth.make_station_db(w, 1500, synth.tile_bbox(buf=3.0), f, days, seed=99) for w in (0, 1)]
    pairs = [[TwxiContext(d, np.isnan(d.stns[db.BAD])) for d in dd] for dd in (env["da"], da_b)]
    wrks = [synth.make_wrk_chk(f, synth.TILE_ROW0 + 20, synth.TILE_COL0 + 30, 40, 40),
            synth.make_wrk_chk(f, synth.TILE_ROW0 + 150, synth.TILE_COL0 + 100, 40, 40)]
    ref = [interp_chunk(p[0], p[1], w) for p, w in zip(pairs, wrks)]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for p, s in zip(pairs, streams):
        for c in p:
            c.set_stream(s.cuda_stream)
    wd = [torch.from_numpy(w).cuda() for w in wrks]
    torch.cuda.synchronize()
    outs = [None, None]
    for rep in range(3):                                             # interleaved submissions, nothing waits in between
        for i in (0, 1):
            outs[i] = interp_chunk(pairs[i][0], pairs[i][1], wd[i], out=outs[i])
    torch.cuda.synchronize()
    for i in (0, 1):
        for k, v in ref[i].items():
            np.testing.assert_array_equal(np.asarray(v), outs[i][k].cpu().numpy(), err_msg="pair %d %s" % (i, k))


def test_interp_pt_with_stns_rm(env):
    """PtInterpTair.interp_pt(stns_rm=...) (interp_tair.py:526-592): the station id is mapped to each variable's own index."""
    from topowx_b200.interp import PtInterpTair
    db = env["db"]
    # make the two station tables differ so that one id has different indices in the two contexts
    da0, da1 = env["da"]
    pti = PtInterpTair(da0, da1)
    opti = o.PtInterpTair(env["oda"][0], env["oda"][1])
    both = np.intersect1d(da0.stn_ids[np.isnan(da0.stns[db.BAD])], da1.stn_ids[np.isnan(da1.stns[db.BAD])])
    sid = next(x for x in both[200:] if pti.ctx_tmin.rm_indices(x)[0, 0] != pti.ctx_tmax.rm_indices(x)[0, 0])
    s = da0.stns[da0.stn_idxs[sid]]
    synth, f = env["synth"], env["f"]
    lat, lon = float(s[db.LAT]) + 0.003, float(s[db.LON]) - 0.002
    elev = float(f.elev(lon, lat))
    for pt in (pti.a_pt, opti.a_pt):
        pt[db.LAT], pt[db.LON], pt[db.ELEV], pt[db.TDI] = lat, lon, elev, float(f.tdi(lon, lat))
        pt[db.CLIMDIV] = float(f.climdiv(lon, lat))
        for m in range(1, 13):
            pt["tmin%02d" % m] = float(f.lst(0, m, lon, lat, elev))
            pt["tmax%02d" % m] = float(f.lst(1, m, lon, lat, elev))
    r, orr = pti.interp_pt(stns_rm=sid), opti.interp_pt(stns_rm=sid)
    for a, b in zip(r[:6], orr[:6]):
        assert np.abs(np.asarray(a) - np.asarray(b)).max() < TOL_C
    r0 = pti.interp_pt()
    assert np.abs(r0[2] - r[2]).max() > 1e-6                         # leaving the nearest station out changes the normals


def test_neighbour_count_above_kriging_limit_is_per_point(env):
    """krig(nnghs=...) above TWXI_MAX_KRIG_NNGHS: that point reports TWXI_ST_LIMIT, the others of the batch are computed."""
    from topowx_b200 import _lib
    from topowx_b200.context import TwxiContext
    db = env["db"]
    da = env["da"][0]
    ctx = TwxiContext(da, np.isnan(da.stns[db.BAD]))
    synth, f = env["synth"], env["f"]
    lat = np.array([40.3, 40.6, 40.9])
    lon = np.array([-99.2, -98.8, -98.5])
    elev = f.elev(lon, lat)
    lst = np.stack([f.lst(0, m, lon, lat, elev) for m in range(1, 13)], axis=1)
    mean, var, st = ctx.krig(lat, lon, elev, lst, mth=4, nnghs=np.array([60, 200, 90]))
    assert st.tolist() == [0, _lib.ST_LIMIT, 0]
    assert np.isfinite(mean[[0, 2], 0]).all() and mean[1, 0] == _lib.FILL_F8


def test_variogram_fit_and_krig_all_vs_oracle(env):
    """twxi_fit_vario / twxi_krig_all (BuildKrigParams, KrigTairAll, XvalTairNorm, StationKrigParams)
    against the oracle's restatement of R get_vario_params / krig_all."""
    from topowx_b200.interp import XvalTairNorm, StationKrigParams
    db = env["db"]
    da, oda = env["da"][1], env["oda"][1]
    good = da.stn_ids[np.isnan(da.stns[db.BAD]) & np.isfinite(da.stns[db.MASK])]
    sids = good[[40, 300, 801, 1200]]
    abw = [35, 57, 92, 147]
    xv, oxv = XvalTairNorm(da, "tmax"), o.XvalTairNorm(oda)
    err, st = xv.run_xval_batch(sids, abw)
    assert err.shape == (4, 12, 4) and np.all(st == 0)
    worst = 0.0
    for i, sid in enumerate(sids):
        oerr = oxv.run_xval(sid, abw)
        worst = max(worst, np.abs(oerr - err[i]).max())
        assert np.abs(oerr - err[i]).max() < TOL_C, (sid, np.abs(oerr - err[i]).max())
    e1 = xv.run_xval(sids[1], np.array(abw))
    assert np.array_equal(e1, err[1])
    print("XvalTairNorm worst |d err| = %.3e C" % worst)
    # step 22: parameters at the stations themselves (the station is its own neighbour at distance 0)
    kp = StationKrigParams(da, "tmax")
    obk = o.BuildKrigParams(o.StationSelect(oda, np.isnan(da.stns[db.BAD])))
    v, st = kp.get_krig_params_batch(sids)
    assert np.all(st == 0)
    for i, sid in enumerate(sids):
        pt = da.stns[da.stn_idxs[sid]]
        for m in (1, 6, 11):
            nug, psill, rng = obk.get_krig_params(pt, m)
            got = v[i, m - 1]
            assert abs(got[0] - nug) <= 1e-7 * max(abs(nug), 1e-3), (sid, m, got, (nug, psill, rng))
            assert abs(got[1] - psill) <= 1e-7 * max(abs(psill), 1e-3), (sid, m, got, (nug, psill, rng))
            assert abs(got[2] - rng) <= 1e-5 * max(abs(rng), 1e-3), (sid, m, got, (nug, psill, rng))
    nugs, psills, rngs = kp.get_krig_params(sids[2])
    assert np.array_equal(nugs, v[2, :, 0]) and np.array_equal(rngs, v[2, :, 2])


def test_interp_to_lonlat_from_predictor_rasters(env, tmp_path):
    """PtInterpTair.interp_to_lonlat (interp_tair.py:513-524): predictors looked up in the rasters of a PredictorStore; the
    result equals the chunk result of the cell that contains the point, and a masked cell raises like the reference."""
    from topowx_b200.context import interp_chunk
    from topowx_b200.interp import PtInterpTair, PredictorStore
    synth, db = env["synth"], env["db"]
    r0, c0 = synth.TILE_ROW0 + 100, synth.TILE_COL0 + 100
    store = PredictorStore.create_synthetic(str(tmp_path / "rasters"), env["f"], r0, c0, 12, 16)
    pti = PtInterpTair(env["da"][0], env["da"][1], aux_fpaths=store)
    wrk = synth.make_wrk_chk_grid(env["f"], r0, c0, 12, 16)
    wrk[2, 3, 4] = 0                                                   # one cell taken out of the chunk's mask
    out = interp_chunk(pti.ctx_tmin, pti.ctx_tmax, wrk)
    lon, lat = synth.grid_lons(c0 + 9) + 0.002, synth.grid_lats(r0 + 5) - 0.003
    tmin_dly, tmax_dly, tmin_norms, tmax_norms, tmin_se, tmax_se, ninvalid = pti.interp_to_lonlat(lon, lat)
    assert pti.a_pt[db.LON] == synth.grid_lons(c0 + 9) and pti.a_pt[db.LAT] == synth.grid_lats(r0 + 5)
    assert np.abs(tmin_norms - out["tmin_norm"][:, 5, 9]).max() < 1e-6
    assert np.abs(tmax_se - out["tmax_se"][:, 5, 9]).max() < 1e-6
    scale = 0.01
    assert np.abs(tmin_dly - out["tmin"][:, 5, 9] * scale).max() <= 0.0051
    assert ninvalid == out["ninvalid"][5, 9]
    mask = np.array(pti.pGrids.rasters["mask"])                        # the store's rasters are read-only maps
    mask[3, 4] = 0
    pti.pGrids.rasters["mask"] = mask
    with pytest.raises(Exception, match="outside interpolation region"):
        pti.interp_to_lonlat(synth.grid_lons(c0 + 4), synth.grid_lats(r0 + 3))
