"""Feature / match stores with the h5py surface the reference uses:

    fd = open_store(path, 'a');  grp = fd.create_group(name);  grp.create_dataset(key, data=array)
    fd[name][key][()]  /  fd[name][key].__array__()  /  name in fd  /  del fd[name]  /  fd.keys()  /  grp.items()  /  fd.close()

(reference ``localization/extract_features.py:223-238``, ``match_features_batch.py:90-129``, ``pose_estimator.py:92-106``).
``open_store`` returns a real ``h5py.File`` when h5py is installed -- the files are then the reference's own format,
byte-compatible with its readers.  h5py is not part of this image, so the fallback is ``ArrayStore``: the same group /
dataset interface over ONE ``.npz`` archive per store (zip of ``.npy`` members named ``group/key``), loaded lazily,
written on ``close()``.  Layouts and dtypes of the datasets are the reference's: features ``keypoints`` float64 [N,2],
``descriptors`` [D,N], ``scores`` [N], ``image_size`` [2]; matches ``matches0`` int16 [N], ``matching_scores0`` float16 [N].
"""
from __future__ import annotations

import os
import zipfile
from pathlib import Path
from typing import Dict, Iterator, Optional

import numpy as np


class _Dataset:
    """A stored array with h5py's read accessors (``ds[()]``, ``ds[...]``, ``np.asarray(ds)``, ``ds.shape``)."""

    def __init__(self, loader):
        self._loader, self._value = loader, None

    def _get(self) -> np.ndarray:
        if self._value is None:
            self._value = np.asarray(self._loader())
        return self._value

    def __getitem__(self, idx):
        v = self._get()
        return v if (isinstance(idx, tuple) and len(idx) == 0) else v[idx]

    def __array__(self, dtype=None, copy=None):
        v = self._get()
        return v.astype(dtype) if dtype is not None else v

    def __iter__(self):
        return iter(self._get())

    def __len__(self):
        return len(self._get())

    @property
    def shape(self):
        return self._get().shape

    @property
    def dtype(self):
        return self._get().dtype


class _Group:
    def __init__(self, store: 'ArrayStore', name: str):
        self._store, self.name = store, name
        self._data: Dict[str, _Dataset] = {}

    def create_dataset(self, key: str, data=None, **kwargs) -> _Dataset:
        if key in self._data:
            raise ValueError(f'dataset {self.name}/{key} already exists')
        arr = np.asarray(data)
        self._data[key] = _Dataset(lambda a=arr: a)
        self._store._dirty = True
        return self._data[key]

    def __getitem__(self, key: str) -> _Dataset:
        return self._data[key]

    def __contains__(self, key) -> bool:
        return key in self._data

    def keys(self):
        return self._data.keys()

    def items(self):
        return self._data.items()

    def __iter__(self):
        return iter(self._data)


class ArrayStore:
    """h5py.File look-alike over one ``.npz`` archive.  Modes: 'r' (must exist), 'a' (read / append), 'w' (truncate)."""

    def __init__(self, path, mode: str = 'r', **kwargs):
        self.path, self.mode = Path(path), mode
        self._groups: Dict[str, _Group] = {}
        self._dirty, self._zip = False, None
        if mode not in ('r', 'a', 'w', 'r+'):
            raise ValueError(mode)
        if self.path.exists() and mode != 'w':
            self._zip = zipfile.ZipFile(str(self.path), 'r')
            for member in self._zip.namelist():
                stem = member[:-4] if member.endswith('.npy') else member
                gname, _, key = stem.rpartition('/')
                grp = self._groups.setdefault(gname, _Group(self, gname))
                grp._data[key] = _Dataset(lambda m=member: self._load(m))
        elif mode in ('r', 'r+'):
            raise FileNotFoundError(str(self.path))

    def _load(self, member: str) -> np.ndarray:
        with self._zip.open(member) as f:
            return np.lib.format.read_array(f, allow_pickle=False)

    # -- h5py.File surface ---------------------------------------------------------------------------------------
    def create_group(self, name: str) -> _Group:
        if self.mode == 'r':
            raise OSError('store opened read-only')
        if name in self._groups:
            raise ValueError(f'group {name} already exists')
        self._groups[name] = _Group(self, name)
        self._dirty = True
        return self._groups[name]

    def __getitem__(self, name: str) -> _Group:
        return self._groups[name]

    def __delitem__(self, name: str):
        del self._groups[name]
        self._dirty = True

    def __contains__(self, name) -> bool:
        return name in self._groups

    def keys(self):
        return self._groups.keys()

    def __iter__(self) -> Iterator[str]:
        return iter(self._groups)

    def __len__(self):
        return len(self._groups)

    def flush(self):
        if not self._dirty or self.mode == 'r':
            return
        arrays = {f'{g}/{k}': np.asarray(ds) for g, grp in self._groups.items() for k, ds in grp.items()}
        if self._zip is not None:
            self._zip.close()
            self._zip = None
        tmp = self.path.with_suffix(self.path.suffix + '.tmp')
        self.path.parent.mkdir(parents=True, exist_ok=True)
        with zipfile.ZipFile(str(tmp), 'w', zipfile.ZIP_STORED) as z:
            for name, arr in arrays.items():
                with z.open(name + '.npy', 'w', force_zip64=True) as f:
                    np.lib.format.write_array(f, arr, allow_pickle=False)
        os.replace(tmp, self.path)
        for g, grp in self._groups.items():  # keep serving reads from memory
            for k in list(grp._data):
                grp._data[k] = _Dataset(lambda a=arrays[f'{g}/{k}']: a)
        self._dirty = False

    def close(self):
        self.flush()
        if self._zip is not None:
            self._zip.close()
            self._zip = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def have_h5py() -> bool:
    try:
        import h5py  # noqa: F401
        return True
    except ImportError:
        return False


def open_store(path, mode: str = 'r', backend: Optional[str] = None, **kwargs):
    """``h5py.File(path, mode)`` when h5py is importable (or ``backend='h5py'``), else ``ArrayStore``.  A file written by one
    backend is recognised by its magic bytes, so a store is always re-opened with the backend that wrote it."""
    p = Path(path)
    if backend is None and p.exists() and mode != 'w':
        with open(p, 'rb') as f:
            magic = f.read(8)
        backend = 'h5py' if magic.startswith(b'\x89HDF') else 'npz'
    if backend is None:
        backend = 'h5py' if have_h5py() else 'npz'
    if backend == 'h5py':
        import h5py
        return h5py.File(str(p), mode, **kwargs)
    kwargs.pop('libver', None)
    return ArrayStore(p, mode)
