"""Solution storage with the reference's indexing grammar.

The reference stores trajectories in ``Series``/``TimeSeries`` objects and reads them back with ids such as
``'x'``, ``'x:f'`` (final column), ``'x:0'`` (first column), ``'x_1'`` (one row by name)
(``hilo_mpc/modules/base.py:2426-2470``).  This is the API-compatibility subset of that container: columns are
time points; values are NumPy arrays (single instance) — batched runs keep their data in torch tensors on the
device and only store what the user asked for.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np


class TimeSeries:
    def __init__(self, **groups: Sequence[str]):
        self._names: Dict[str, List[str]] = {}
        self._data: Dict[str, Optional[np.ndarray]] = {}
        for key, names in groups.items():
            self.setup(key, names)

    def setup(self, key: str, names: Sequence[str]) -> None:
        self._names[key] = list(names)
        self._data[key] = None

    def is_empty(self, key: Optional[str] = None) -> bool:
        if key is None:
            return all(v is None or v.shape[1] == 0 for v in self._data.values())
        v = self._data.get(key)
        return v is None or v.shape[1] == 0

    def keys(self):
        return self._data.keys()

    def get_names(self, key: str) -> List[str]:
        return list(self._names[key])

    def add(self, key: str, value) -> None:
        value = np.asarray(value, dtype=float)
        if value.ndim == 0:
            value = value.reshape(1, 1)
        elif value.ndim == 1:
            value = value.reshape(-1, 1)
        if self._data.get(key) is None:
            self._data[key] = value.copy()
        else:
            self._data[key] = np.concatenate([self._data[key], value], axis=1)

    def set(self, key: str, value) -> None:
        self._data[key] = None
        self.add(key, value)

    def update(self, **kwargs) -> None:
        for key, value in kwargs.items():
            if value is not None:
                self.add(key, value)

    def get_by_id(self, ident: str) -> np.ndarray:
        key, _, sel = ident.partition(':')
        if key in self._data:
            data = self._data[key]
            if data is None:
                data = np.zeros((len(self._names.get(key, [])), 0))
        else:
            data = None
            for grp, names in self._names.items():
                if key in names:
                    full = self._data[grp]
                    if full is None:
                        full = np.zeros((len(names), 0))
                    data = full[names.index(key):names.index(key) + 1]
                    break
            if data is None:
                raise KeyError(ident)
        if sel == '':
            return data
        if sel == 'f':
            return data[:, -1:]
        if sel == '0':
            return data[:, :1]
        return data[:, int(sel):int(sel) + 1]

    def __getitem__(self, ident: str) -> np.ndarray:
        return self.get_by_id(ident)

    def __contains__(self, key: str) -> bool:
        return key in self._data

    def reset(self, keep_first: bool = False) -> None:
        for key, v in self._data.items():
            if v is not None:
                self._data[key] = v[:, :1].copy() if keep_first else None

    def to_dict(self) -> Dict[str, np.ndarray]:
        return {k: (None if v is None else v.copy()) for k, v in self._data.items()}
