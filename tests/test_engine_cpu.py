"""Host logic of the decoder step that needs no GPU: phase 1 (`DecoderEngine._plan`) must refuse a bad step before any
stream state (length, pages, free list) has changed."""
import pytest
import torch

from mmduet_b200 import _lib
from mmduet_b200.config import ModelConfig
from mmduet_b200.engine import PAGE, DecoderEngine, KVStorage


def _bare_engine(n_pages=8, max_context=4096):
    eng = DecoderEngine.__new__(DecoderEngine)
    eng.cfg = ModelConfig()
    eng.max_context = max_context
    eng.n_pages = n_pages
    eng._free = list(range(n_pages - 1, -1, -1))
    return eng


def _stream(eng, length):
    st = KVStorage(eng)
    st.ensure(length)
    st.length = length
    return st


def _snapshot(eng, streams):
    return (list(eng._free), [(s.length, list(s.pages)) for s in streams])


def test_plan_refuses_before_mutating():
    eng = _bare_engine()
    H = eng.cfg.hidden
    a, b = _stream(eng, 100), _stream(eng, 70)
    snap = _snapshot(eng, [a, b])
    ok = dict(storage=a, past=64, ids=[1, 2, 3])                     # a rollback to 64 + 3 tokens: valid by itself
    bad_items = [
        dict(storage=b, past=71, ids=[1]),                           # stale view (longer than the stream)
        dict(storage=b, past=70, ids=[eng.cfg.vocab]),               # token id out of range
        dict(storage=b, past=70, ids=[]),                            # empty item
        dict(storage=b, past=70, embeds=torch.zeros(3, H + 8)),      # wrong hidden size
        dict(storage=b, past=70, ids=[1] * (eng.max_context)),       # context limit
        dict(storage=b, past=70, ids=[1] * (PAGE * 7)),              # KV pool exhausted (8 pages, 4 in use, 7 more needed)
        dict(storage=a, past=64, ids=[5]),                           # same stream twice
    ]
    for bad in bad_items:
        with pytest.raises(_lib.MmdError):
            eng._plan([ok, bad], "last", "none")
        assert _snapshot(eng, [a, b]) == snap, bad
    with pytest.raises(_lib.MmdError):
        eng._plan([dict(storage=b, past=0, ids=[1, 2], score_rows=[2])], "frame_ends", "none")
    assert _snapshot(eng, [a, b]) == snap
    plan = eng._plan([ok, dict(storage=b, past=70, embeds=torch.zeros(5, H))], "last", "none")
    assert [(p[2], p[5], p[6]) for p in plan] == [(64, 3, 67), (70, 5, 75)]
    assert plan[1][3] == [-1, -2, -3, -4, -5]
    assert _snapshot(eng, [a, b]) == snap                            # planning itself never mutates


def test_plan_counts_pages_recycled_by_rollbacks():
    eng = _bare_engine(n_pages=6)
    a, b = _stream(eng, 4 * PAGE), _stream(eng, 2 * PAGE)            # pool full: 6 of 6 pages in use
    # a rolls back to one page, b grows by two: fits only because a's rollback recycles pages
    plan = eng._plan([dict(storage=b, past=2 * PAGE, ids=[1] * (2 * PAGE)), dict(storage=a, past=PAGE, ids=[1])], "last", "none")
    assert len(plan) == 2
    with pytest.raises(_lib.MmdError):
        eng._plan([dict(storage=b, past=2 * PAGE, ids=[1] * (2 * PAGE)), dict(storage=a, past=3 * PAGE + 1, ids=[1])], "last", "none")
