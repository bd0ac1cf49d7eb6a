"""Gallery persistence formats on either side of the hot path.

* the reference stores one Mongo document per face, ``{"target": name, "embedding": s}``,
  where ``s`` is ``json.dumps(list_of_floats)`` or, with encryption enabled,
  ``base64(Fernet(key).encrypt(json.dumps(...)))`` (backend/app/utils/db.py:238-267,
  460-471); ``retrieve_all_embeddings`` (db.py:484-490) returns those documents but has no
  caller — the reference never reloads its gallery.  ``load_faces`` is that missing
  start-up load: decode every document and put the whole gallery into HBM in one batch.
* ``save_snapshot`` / ``load_snapshot``: flat binary snapshot of the HBM gallery
  (``frp_gallery_save``: header + fp32 master rows + tombstones) plus a JSON side-car with
  the names, for galleries too large for per-face documents.
"""
from __future__ import annotations

import base64
import json
import os
from typing import Any, Dict, Iterable, List, Optional

import numpy as np

from .gallery import Gallery, GalleryDict


def _fernet(key: Optional[bytes]):
    if key is None:
        return None
    from cryptography.fernet import Fernet
    return Fernet(key)


def encode_embedding(embedding: List[float], key: Optional[bytes] = None) -> str:
    """db.py:238-251 ``encrypt_embedding``: plaintext JSON without a key, else
    base64(Fernet token of the JSON)."""
    raw = json.dumps(list(embedding))
    f = _fernet(key)
    if f is None:
        return raw
    return base64.b64encode(f.encrypt(raw.encode("utf-8"))).decode("utf-8")


def decode_embedding(stored: str, key: Optional[bytes] = None) -> List[float]:
    """db.py:254-267 ``decrypt_embedding``; returns [] on any error like the reference."""
    try:
        f = _fernet(key)
        if f is None:
            return json.loads(stored)
        return json.loads(f.decrypt(base64.b64decode(stored)).decode("utf-8"))
    except Exception:
        return []


def load_faces(encodings: GalleryDict, docs: Iterable[Dict[str, Any]], key: Optional[bytes] = None) -> int:
    """Bulk-load ``retrieve_all_embeddings()``-style documents into the gallery (document
    order = insertion order = tie order).  Documents that fail to decode are skipped, as the
    reference's ``decrypt_embedding`` yields ``[]`` for them.  A later document for a name
    already loaded replaces the value in place (dict assignment semantics).  Returns the
    number of faces loaded."""
    names: List[str] = []
    rows: List[List[float]] = []
    pos: Dict[str, int] = {}
    for d in docs:
        name, stored = d.get("target"), d.get("embedding")
        if name is None or stored is None:
            continue
        vec = decode_embedding(stored, key) if isinstance(stored, str) else list(stored)
        if not vec:
            continue
        if name in pos:
            rows[pos[name]] = vec
        elif name in encodings:
            encodings[name] = vec
        else:
            pos[name] = len(names)
            names.append(name)
            rows.append(vec)
    if names:
        dim = len(rows[0])
        good = [i for i, r in enumerate(rows) if len(r) == dim]
        encodings.bulk_load([names[i] for i in good], np.asarray([rows[i] for i in good], dtype=np.float64))
        return len(good)
    return 0


def save_snapshot(encodings: GalleryDict, path: str) -> None:
    """Compact the gallery, write ``path`` (binary rows) and ``path + '.names.json'``."""
    encodings.compact()
    if encodings.gallery is None:
        raise ValueError("empty gallery")
    encodings.gallery.save(path)
    tmp = path + ".names.json.tmp"
    with open(tmp, "w") as f:
        json.dump({"metric": encodings.metric, "names": encodings.keys()}, f)
    os.replace(tmp, path + ".names.json")      # atomic like the reference's backups (face_service.py:731-741)


def load_snapshot(path: str, device: int = 0, keep_host: bool = False) -> GalleryDict:
    with open(path + ".names.json") as f:
        meta = json.load(f)
    g = Gallery.load(path, device)
    rows, live = g.counts()
    if rows != live or rows != len(meta["names"]):
        raise ValueError("snapshot and name list disagree (%d rows, %d live, %d names)" % (rows, live, len(meta["names"])))
    d = GalleryDict(metric=g.metric, device=device, keep_host=keep_host)
    d._gallery = g
    d._names = list(meta["names"])
    d._rows = {n: i for i, n in enumerate(d._names)}
    return d
