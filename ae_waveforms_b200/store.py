"""On-disk side of the path: per-chunk shard files with a resume manifest, and the three final arrays.

Replaces, for the batched path, the reference's per-case parquet files and their "already done" check
(data_extract_vitaldb.py:37-42, 173-179), the ``data_stats.pkl`` bookkeeping (:206-224: case ids that
produced rows and the total row count) and ``data_create_arrays.py:17-58`` (stack per case in case order,
``np_waves`` float32, ``np_features`` / ``np_info`` float64, saved as ``f"{path}np_waves"`` ...).
The object-cell parquet schema itself is not reproduced: its 2-D ``ecg_samples_aligned`` cells cannot even be
written by the pyarrow of this image; the rows of the three arrays are what every later
stage of the reference reads.

Layout under ``root``:
    shards/<first_case>_<last_case>.npz   one file per finished chunk (waves, features, info, case_ids)
    manifest.json                         {"chunks": [{"file", "case_ids", "rows"}], "version": 1}
    np_waves.npy, np_features.npy, np_info.npy, data_stats.json   written by ``finalize``
A chunk is committed by writing its shard to a temporary name, renaming it, and then rewriting the manifest the
same way, so a killed run loses at most the chunk in flight; ``pending`` returns the case ids still to do.
"""
from __future__ import annotations

import json
import os

import numpy as np

VERSION = 1


class ArrayStore:
    def __init__(self, root: str):
        self.root = root
        self.shard_dir = os.path.join(root, "shards")
        os.makedirs(self.shard_dir, exist_ok=True)
        self.manifest_path = os.path.join(root, "manifest.json")
        self.manifest = {"version": VERSION, "chunks": []}
        if os.path.exists(self.manifest_path):
            with open(self.manifest_path) as f:
                self.manifest = json.load(f)
            if self.manifest.get("version") != VERSION:
                raise ValueError(f"{self.manifest_path}: unknown manifest version")
            # a shard named in the manifest must exist (the manifest is written after the shard)
            self.manifest["chunks"] = [c for c in self.manifest["chunks"]
                                       if os.path.exists(os.path.join(self.shard_dir, c["file"]))]

    # ------------------------------------------------------------------ resume
    def done_case_ids(self) -> set:
        return {int(c) for ch in self.manifest["chunks"] for c in ch["case_ids"]}

    def pending(self, case_ids) -> list:
        """The case ids that still have to be processed (data_extract_vitaldb.py:37-42: a case whose file
        exists is not processed again)."""
        done = self.done_case_ids()
        return [int(c) for c in case_ids if int(c) not in done]

    # ------------------------------------------------------------------ write
    @staticmethod
    def _host(a, dtype):
        if hasattr(a, "detach"):
            a = a.detach().cpu().numpy()
        return np.ascontiguousarray(a, dtype=dtype)

    def add_chunk(self, case_ids, np_waves, np_info, np_features=None) -> str:
        """Commit the rows of one processed chunk of cases (tensors or arrays; cases without rows included in
        ``case_ids`` so that they are not processed again)."""
        case_ids = [int(c) for c in case_ids]
        if not case_ids:
            return ""
        overlap = self.done_case_ids().intersection(case_ids)
        if overlap:
            raise ValueError(f"cases already stored: {sorted(overlap)[:5]}")
        waves = self._host(np_waves, np.float32)
        info = self._host(np_info, np.float64).reshape(-1, 2)
        feats = np.full((len(waves), 35), np.nan) if np_features is None else self._host(np_features, np.float64)
        if not (len(waves) == len(info) == len(feats)) or (len(waves) and waves.shape[1] != 320) or feats.shape[1] != 35:
            raise ValueError("row counts / shapes of the three arrays do not agree")
        if len(info) and not set(np.unique(info[:, 0]).astype(np.int64)).issubset(case_ids):
            raise ValueError("np_info names a case that is not in case_ids")
        name = f"{min(case_ids)}_{max(case_ids)}.npz"
        tmp = os.path.join(self.shard_dir, name + ".tmp")
        with open(tmp, "wb") as f:
            np.savez(f, waves=waves, features=feats, info=info, case_ids=np.asarray(case_ids, dtype=np.int64))
        os.replace(tmp, os.path.join(self.shard_dir, name))
        self.manifest["chunks"].append({"file": name, "case_ids": case_ids, "rows": int(len(waves))})
        tmpm = self.manifest_path + ".tmp"
        with open(tmpm, "w") as f:
            json.dump(self.manifest, f)
        os.replace(tmpm, self.manifest_path)
        return name

    # ------------------------------------------------------------------ final arrays
    def finalize(self, prefix: str | None = None) -> dict:
        """data_create_arrays.py:17-58: the rows of all cases in ascending case order (rows of a case in idx_end
        order), ``np_waves`` float32, ``np_features`` and ``np_info`` float64, saved with the reference's file
        names; plus the reference's data_stats (case ids with rows, total rows)."""
        parts = []
        for ch in self.manifest["chunks"]:
            with np.load(os.path.join(self.shard_dir, ch["file"])) as z:
                parts.append((z["waves"], z["features"], z["info"]))
        waves = np.concatenate([p[0] for p in parts]) if parts else np.zeros((0, 320), np.float32)
        feats = np.concatenate([p[1] for p in parts]) if parts else np.zeros((0, 35))
        info = np.concatenate([p[2] for p in parts]) if parts else np.zeros((0, 2))
        order = np.lexsort((info[:, 1], info[:, 0])) if len(info) else np.zeros(0, dtype=np.int64)
        waves, feats, info = waves[order].astype(np.float32), feats[order].astype(np.float64), info[order]
        prefix = os.path.join(self.root, "") if prefix is None else prefix
        np.save(f"{prefix}np_waves", waves)
        np.save(f"{prefix}np_features", feats)
        np.save(f"{prefix}np_info", info)
        ids, counts = np.unique(info[:, 0].astype(np.int64), return_counts=True) if len(info) else ([], [])
        stats = {"case_ids": [int(c) for c in ids], "samples_per_case_id": [int(c) for c in counts],
                 "samples_total": int(len(info))}                    # data_extract_vitaldb.py:216
        with open(f"{prefix}data_stats.json", "w") as f:
            json.dump(stats, f)
        return dict(np_waves=waves, np_features=feats, np_info=info, **stats)


def process_to_store(store: ArrayStore, pre, records_of, case_ids, chunk_cases: int = 100, features: bool = True):
    """Resumable driver: ``records_of(ids)`` returns the 1-D float64 tracks of the given case ids (the
    loader, load_record.py:7-37, stays with the caller).  Cases already in the store are skipped."""
    from .batched import RecordBatch
    todo = store.pending(case_ids)
    for c0 in range(0, len(todo), chunk_cases):
        ids = todo[c0:c0 + chunk_cases]
        res = pre.run(RecordBatch.from_host(records_of(ids), ids, device=pre.device), features=features)
        store.add_chunk(ids, res["np_waves"], res["np_info"], res.get("np_features"))
    return store
