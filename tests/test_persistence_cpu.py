"""Persistence formats (row N3) on CPU: the document encoding of
backend/app/utils/db.py:238-267 and the start-up bulk load, checked against the
reference's own encrypt/decrypt functions when /root/reference is present."""
import numpy as np
import pytest

import ref_loader
from frp_b200.gallery import GalleryDict
from frp_b200.persistence import decode_embedding, encode_embedding, load_faces
from test_capi_cpu import _HostGallery

Fernet = pytest.importorskip("cryptography.fernet").Fernet


def test_document_roundtrip_plain_and_encrypted():
    v = np.random.default_rng(0).normal(size=128).tolist()
    assert decode_embedding(encode_embedding(v)) == v
    key = Fernet.generate_key()
    s = encode_embedding(v, key)
    assert s != encode_embedding(v) and decode_embedding(s, key) == v
    assert decode_embedding(s, Fernet.generate_key()) == []        # wrong key -> [] like the reference
    assert decode_embedding("not json") == [] and decode_embedding("???", key) == []


@pytest.mark.skipif(not ref_loader.reference_available(), reason="/root/reference not present")
def test_document_format_matches_the_reference():
    _, _, db = ref_loader.load_reference()
    v = np.random.default_rng(1).normal(size=128).tolist()
    assert db.encrypt_embedding(v) == encode_embedding(v)          # DISABLE_ENCRYPTION=true: plain JSON
    assert db.decrypt_embedding(encode_embedding(v)) == v
    key = Fernet.generate_key()
    old = (db.fernet, db.DISABLE_ENCRYPTION)
    try:
        db.fernet, db.DISABLE_ENCRYPTION = Fernet(key), False
        assert decode_embedding(db.encrypt_embedding(v), key) == v  # reference writes, we read
        assert db.decrypt_embedding(encode_embedding(v, key)) == v  # we write, reference reads
    finally:
        db.fernet, db.DISABLE_ENCRYPTION = old


def test_bulk_load_of_documents_keeps_dict_semantics():
    rng = np.random.default_rng(2)
    key = Fernet.generate_key()
    vecs = {n: rng.normal(size=16).tolist() for n in ("a", "b", "c", "d")}
    docs = [{"target": "a", "embedding": encode_embedding(vecs["a"], key)},
            {"target": "b", "embedding": encode_embedding(vecs["b"], key)},
            {"target": "broken", "embedding": "garbage"},
            {"target": "c", "embedding": encode_embedding(vecs["c"], key)},
            {"target": "a", "embedding": encode_embedding(vecs["d"], key)},     # later doc for "a" wins, keeps position
            {"target": "short", "embedding": encode_embedding([1.0, 2.0], key)},
            {"embedding": "x"}]
    d = GalleryDict(gallery_factory=_HostGallery)
    assert load_faces(d, docs, key) == 3
    assert d.keys() == ["a", "b", "c"]
    assert np.allclose(d["a"], vecs["d"]) and np.allclose(d["c"], vecs["c"])
    assert load_faces(d, [{"target": "b", "embedding": encode_embedding(vecs["d"], key)}], key) == 0
    assert np.allclose(d["b"], vecs["d"]) and d.keys() == ["a", "b", "c"]
