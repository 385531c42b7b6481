"""CPU: the C-ABI shared library loads and exports every symbol include/cy4.h declares
(no compute calls without a GPU)."""
import os
import re

from conftest import PKG, ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "cy4.h")).read()
    return sorted(set(re.findall(r"CY4_API\s+[\w\s\*]+?\b(cy4_\w+)\s*\(", text)))


def test_header_symbols_exported():
    import __graft_entry__ as ge
    ge.build()
    from cy4 import _lib
    L = _lib.lib()
    names = _declared()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert L.cy4_version() == 100
    assert L.cy4_last_error() is not None


def test_ctypes_signatures_cover_header():
    from cy4 import _lib, _sigs_engine
    bound = set(_lib._SIGS) | set(_sigs_engine.SIGS)
    assert set(_declared()) <= bound, set(_declared()) - bound


def test_no_fallback_without_library(tmp_path, monkeypatch):
    """The product path must fail loudly when the CUDA library is missing."""
    import pytest
    from cy4 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "SO_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError):
        _lib.lib()


def test_product_never_imports_oracle():
    bad = []
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, re.M) or "rbox_oracle" in src:
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_set_option_names_and_ranges():
    """cy4_set_option is host-only state: every documented tunable parses, bad names / values are rejected with an
    error string, and the defaults are restored."""
    from cy4 import _lib
    L = _lib.lib()
    ok = [(b"conv_cluster", 2), (b"conv_cluster", 1), (b"wgrad_cluster", 2), (b"wgrad_cluster", 1), (b"tma_store", 0), (b"tma_store", 1),
          (b"kblocks_per_slot", 1), (b"kblocks_per_slot", 4), (b"conv_pair", 0), (b"conv_pair", 1), (b"conv1x1_matrix", 1), (b"conv1x1_matrix", 0),
          (b"slab_stats", 0), (b"slab_stats", 1), (b"dgrad_interleave", 0), (b"dgrad_interleave", 1), (b"ew_carveout", 1), (b"ew_carveout", 0), (b"pdl", 1), (b"pdl", 0), (b"ew_fwd_blocks_per_sm", 6), (b"ew_fwd_blocks_per_sm", 3), (b"ew_bwd_blocks_per_sm", 6), (b"ew_bwd_blocks_per_sm", 2), (b"debug", 0)]
    for name, value in ok:
        assert L.cy4_set_option(name, value) == 0, (name, value)
    for name, value in [(b"no_such_option", 1), (b"wgrad_cluster", 3), (b"conv_cluster", 3)]:
        assert L.cy4_set_option(name, value) < 0, (name, value)
        assert L.cy4_last_error()
