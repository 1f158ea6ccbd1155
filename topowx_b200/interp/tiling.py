'''
Tiling of the interpolation grid: the `Tiler` / `TileGridInfo` / `TileWriter` interfaces of
twx/interp/tiling.py:44-537.  `Tiler` works on in-memory arrays (anything indexable like the reference's
netCDF4 variables); `partition_chunks` is the multi-GPU replacement for step25's coordinator rank.
netCDF tile writing stays on the reference path: `TileWriter` needs the netCDF4 module and otherwise raises.
'''

__all__ = ['Tiler', 'TileGridInfo', 'TileWriter', 'partition_chunks']

import os

import numpy as np


class Tiler():
    '''
    Breaks an interpolation grid into tiles and work chunks (tiling.py:44-275).
    '''

    def __init__(self, ds_mask, ds_attr_ls, tile_size_y, tile_size_x, chk_size_y, chk_size_x, path_out=None,
                 process_tiles=False):
        '''
        ds_mask : mapping with 'mask' (2-D), 'lon' (1-D), 'lat' (1-D, descending) – a netCDF4.Dataset's
            `.variables` or a dict of arrays.
        ds_attr_ls : list of (name, 2-D array-like) auxiliary predictor grids, in work-chunk plane order.
        '''
        v = ds_mask.variables if hasattr(ds_mask, 'variables') else ds_mask
        self.mask = np.array(v['mask'][:], dtype=bool)
        self.lons = np.asarray(v['lon'][:])
        self.lats = np.asarray(v['lat'][:])
        self.nrows = self.lats.size
        self.ncols = self.lons.size
        if self.nrows % tile_size_y or self.ncols % tile_size_x:
            raise ValueError("the grid size must be evenly divisible by the tile size")
        if tile_size_y % chk_size_y or tile_size_x % chk_size_x:
            raise ValueError("the tile size must be evenly divisible by the chunk size")
        self.tile_size_y = tile_size_y
        self.tile_size_x = tile_size_x
        self.chk_size_y = chk_size_y
        self.chk_size_x = chk_size_x
        self.attrs = []
        for varname, ds in ds_attr_ls:
            attr = ds.variables[varname] if hasattr(ds, 'variables') else ds
            if hasattr(attr, 'set_auto_maskandscale'):
                attr.set_auto_maskandscale(False)
            self.attrs.append(attr)
        self.tile_ids, self.tile_rc = self.__build_tile_dicts(self.nrows, self.ncols, tile_size_y, tile_size_x, self.mask)
        self.chk_size_i = 5 + len(self.attrs)
        self.wrk_chk = np.zeros((self.chk_size_i, chk_size_y, chk_size_x)) * np.nan
        try:
            self.process_tiles = list(process_tiles)
        except TypeError:
            if type(process_tiles) == bool and process_tiles:
                self.process_tiles = self.get_incomplete_tile_nums(path_out)
            else:
                self.process_tiles = None
        self.__set_tile_chks()

    def __set_tile_chks(self):
        k = 0
        self.tile_chks = []
        self.ntiles = 0
        for i in np.arange(0, self.nrows, self.tile_size_y):
            for j in np.arange(0, self.ncols, self.tile_size_x):
                msk_tile = self.mask[i:i + self.tile_size_y, j:j + self.tile_size_x]
                if msk_tile.any():
                    process_tile = self.process_tiles is None or k in self.process_tiles
                    if process_tile:
                        for y in np.arange(0, self.tile_size_y, self.chk_size_y):
                            for x in np.arange(0, self.tile_size_x, self.chk_size_x):
                                self.tile_chks.append((k, int(i), int(j), int(y), int(x)))
                        self.ntiles += 1
                    k += 1
        self.iter_x = 0
        self.ntile_chks = len(self.tile_chks)

    def build_chunk(self, chk, out=None):
        '''The work chunk f8[(5+N), Y, X] of entry `chk` = (k, i, j, y, x) of tile_chks (tiling.py:179-215).'''
        k, i, j, y, x = chk
        w = self.wrk_chk if out is None else out
        cy, cx = self.chk_size_y, self.chk_size_x
        rcgrid = np.mgrid[y:y + cy, x:x + cx]
        w[0, :, :] = rcgrid[0, :, :]
        w[1, :, :] = rcgrid[1, :, :]
        w[2, :, :] = self.mask[i + y:i + y + cy, j + x:j + x + cx]
        w[3, :, :] = self.lats[i + y:i + y + cy][:, None]
        w[4, :, :] = self.lons[j + x:j + x + cx][None, :]
        for z, attr in enumerate(self.attrs):
            w[5 + z, :, :] = attr[i + y:i + y + cy, j + x:j + x + cx]
        return k, w

    def next(self):
        '''
        Next work chunk: (tile number, wrk_chk f8[(5+N), Y, X]); planes 0 row, 1 col, 2 mask, 3 lat, 4 lon,
        then the N predictors.  Raises StopIteration at the end (tiling.py:167-217).
        '''
        if self.iter_x == self.ntile_chks:
            raise StopIteration()
        k, w = self.build_chunk(self.tile_chks[self.iter_x])
        self.iter_x += 1
        return k, w

    __next__ = next

    def __iter__(self):
        return self

    def __build_tile_dicts(self, nrows, ncols, tile_size_y, tile_size_x, mask):
        tile_ids = {}
        tile_info = {}
        x = 0
        cnt_y = 0
        for i in np.arange(0, nrows, tile_size_y):
            cnt_x = 0
            for j in np.arange(0, ncols, tile_size_x):
                if mask[i:i + tile_size_y, j:j + tile_size_x].any():
                    atileid = "".join(["h%02d" % (cnt_x,), "v%02d" % (cnt_y,)])
                    tile_ids[x] = atileid
                    tile_info[atileid] = (int(i), int(j))
                    x += 1
                cnt_x += 1
            cnt_y += 1
        return tile_ids, tile_info

    def build_tile_grid_info(self):
        return TileGridInfo(self.tile_ids, self.tile_rc, self.ntiles, self.lons, self.lats, self.tile_size_y,
                            self.tile_size_x, self.chk_size_y, self.chk_size_x, self.chk_size_i)

    def get_incomplete_tile_nums(self, path_out):
        name_to_id = {a_name: a_id for a_id, a_name in self.tile_ids.items()}
        all_ids = np.unique(list(self.tile_ids.keys()))
        names_done = [n for n in os.listdir(path_out) if n in name_to_id]
        ids_done = np.unique([name_to_id[a_name] for a_name in names_done])
        return all_ids[~np.isin(all_ids, ids_done)]


class TileGridInfo():
    '''Information on tiles and work chunks of an interpolation grid (tiling.py:277-302).'''

    def __init__(self, tile_ids, tile_rc, ntiles, lons, lats, tile_size_y, tile_size_x, chk_size_y, chk_size_x,
                 chk_size_i):
        self.tile_ids = tile_ids
        self.tile_rc = tile_rc
        self.ntiles = ntiles
        self.lons = lons
        self.lats = lats
        self.tile_size_y = tile_size_y
        self.tile_size_x = tile_size_x
        self.chk_size_y = chk_size_y
        self.chk_size_x = chk_size_x
        self.chk_size_i = chk_size_i
        self.chks_per_tile = (tile_size_x // chk_size_x) * (tile_size_y // chk_size_y)
        self.nchks = self.chks_per_tile * ntiles

    def get_tile_id(self, tile_num):
        return self.tile_ids[tile_num]


def partition_chunks(tile_chks, mask, tile_size_y, tile_size_x, world_size, rank=None):
    '''
    Static multi-GPU partition of the ordered chunk list (replaces the coordinator rank's first-come dispatch,
    step25_mpi_interp_tair.py:293-305): whole tiles are assigned to ranks by greedy longest-processing-time on
    the number of unmasked cells, chunks keep their reference order inside a rank.  Deterministic, so every rank
    computes the same partition without communication.  Returns the per-rank lists, or rank's list.
    '''
    tiles = {}
    for c in tile_chks:
        tiles.setdefault(c[0], []).append(c)
    work = []
    for k, chks in tiles.items():
        _, i, j, _, _ = chks[0]
        work.append((int(mask[i:i + tile_size_y, j:j + tile_size_x].sum()), k))
    work.sort(key=lambda t: (-t[0], t[1]))
    loads = [0] * world_size
    owner = {}
    for n, k in work:
        r = min(range(world_size), key=lambda q: (loads[q], q))
        owner[k] = r
        loads[r] += n
    parts = [[c for c in tile_chks if owner[c[0]] == r] for r in range(world_size)]
    return parts if rank is None else parts[rank]


class TileWriter():
    '''
    netCDF tile output (tiling.py:304-537).  netCDF I/O stays on the reference path: this class only exists so
    that drivers written against the reference import cleanly; it needs the netCDF4 module.
    '''

    def __init__(self, tile_grid_info, path_out):
        try:
            import netCDF4  # noqa: F401
        except ImportError:
            raise ImportError("TileWriter writes netCDF tiles and needs the netCDF4 module (not installed); "
                              "use the arrays returned by PtInterpTair.interp_chunk or write .npz tiles")
        raise NotImplementedError("netCDF tile writing stays on the reference path (twx/interp/tiling.py:304-537)")
