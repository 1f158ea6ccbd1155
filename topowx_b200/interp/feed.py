'''
The I/O feed of the gridded interpolation: predictor rasters on disk -> work tiles in pinned host
memory, ahead of the GPU.

The reference's Tiler.next re-reads the 27 predictor windows from netCDF for EVERY 50x50 work chunk
(twx/interp/tiling.py:194-213) on the coordinator rank and ships the chunk over MPI.  Here a tile's windows are read once,
by a background thread, straight into one of a ring of pinned buffers in the work-chunk plane layout
(tiling.py:205-213, step25:273-279), so that reading tile t+1 overlaps the kernels of tile t and the host->device copy
needs no staging copy.  Rasters live in a `PredictorStore`: a directory of .npy files opened as memory maps (netCDF
rasters are accepted when the netCDF4 module imports; raster ingest itself stays on the reference path).
'''

__all__ = ['PredictorStore', 'TileFeed', 'PLANE_NAMES']

import os
import queue
import threading

import numpy as np

# work-chunk planes 5.. in the order of step25:273-279
PLANE_NAMES = ['elev', 'tdi', 'climdiv'] + ['tmin%02d' % m for m in range(1, 13)] + ['tmax%02d' % m for m in range(1, 13)]


class PredictorStore(object):
    '''
    Grid definition (lon, lat, mask) and the 27 auxiliary predictor rasters of an interpolation grid as .npy files in one
    directory: lon.npy [ncols], lat.npy [nrows, descending], mask.npy [nrows, ncols] and one <name>.npy [nrows, ncols]
    per entry of PLANE_NAMES.  `variables` / `attrs()` plug into Tiler the way the netCDF datasets of the reference do
    (step25:266-281).
    '''

    def __init__(self, path):
        self.path = path
        ld = lambda n: np.load(os.path.join(path, n + '.npy'), mmap_mode='r')
        self.variables = {'lon': ld('lon'), 'lat': ld('lat'), 'mask': ld('mask')}
        self.rasters = [ld(n) for n in PLANE_NAMES]
        shp = self.variables['mask'].shape
        for n, r in zip(PLANE_NAMES, self.rasters):
            if r.shape != shp:
                raise ValueError("raster %s has shape %r, the mask %r" % (n, r.shape, shp))

    def attrs(self):
        return list(zip(PLANE_NAMES, self.rasters))

    @staticmethod
    def create_synthetic(path, fields, row0, col0, nrows, ncols, band=250):
        '''Write the rasters of the window [row0, row0+nrows) x [col0, col0+ncols) of the synthetic CONUS grid.'''
        from .. import synth
        os.makedirs(path, exist_ok=True)
        np.save(os.path.join(path, 'lon.npy'), synth.grid_lons(np.arange(col0, col0 + ncols)))
        np.save(os.path.join(path, 'lat.npy'), synth.grid_lats(np.arange(row0, row0 + nrows)))
        out = {n: np.lib.format.open_memmap(os.path.join(path, n + '.npy'), mode='w+', dtype=np.float64, shape=(nrows, ncols))
               for n in PLANE_NAMES}
        mask = np.lib.format.open_memmap(os.path.join(path, 'mask.npy'), mode='w+', dtype=np.int8, shape=(nrows, ncols))
        for r in range(0, nrows, band):
            h = min(band, nrows - r)
            w = synth.make_wrk_chk_grid(fields, row0 + r, col0, h, ncols)
            mask[r:r + h] = w[2] != 0
            for z, n in enumerate(PLANE_NAMES):
                out[n][r:r + h] = w[5 + z]
        for a in list(out.values()) + [mask]:
            a.flush()
        return PredictorStore(path)


class TileFeed(object):
    '''
    Iterator over (tile number, tile id, wrk) for the tiles `tile_nums` of `tiler` (built with chunk size = tile size,
    or any chunk list passed as `chunks`): `wrk` is a pinned float64 torch tensor [5 + N, Y, X] that stays valid until
    `depth - 1` further items have been taken.  A daemon thread reads ahead.
    '''

    def __init__(self, tiler, chunks=None, depth=3, pinned=True):
        import torch
        self.tiler = tiler
        self.chunks = list(tiler.tile_chks if chunks is None else chunks)
        shape = (tiler.chk_size_i, tiler.chk_size_y, tiler.chk_size_x)
        mk = (lambda: torch.empty(shape, dtype=torch.float64).pin_memory()) if pinned else \
             (lambda: torch.empty(shape, dtype=torch.float64))
        self.bufs = [mk() for _ in range(depth)]
        self.free = queue.Queue()
        for i in range(depth):
            self.free.put(i)
        self.ready = queue.Queue(maxsize=depth)
        self.held = []
        self.bytes_read = 0
        self.thread = threading.Thread(target=self._reader, daemon=True)
        self.thread.start()

    def _reader(self):
        try:
            for chk in self.chunks:
                i = self.free.get()
                k, _ = self.tiler.build_chunk(chk, out=self.bufs[i].numpy())
                self.bytes_read += self.bufs[i].numel() * 8
                self.ready.put((k, i))
            self.ready.put(None)
        except Exception as e:                                    # surface reader failures in the consumer
            self.ready.put(e)

    def __iter__(self):
        return self

    def __next__(self):
        item = self.ready.get()
        if item is None:
            raise StopIteration()
        if isinstance(item, Exception):
            raise item
        k, i = item
        self.held.append(i)
        while len(self.held) > len(self.bufs) - 1:                # the oldest buffer goes back to the reader
            self.free.put(self.held.pop(0))
        return k, self.tiler.tile_ids[k], self.bufs[i]
