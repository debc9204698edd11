"""The C-ABI library loads on a CPU-only box and exports exactly what include/ppg_b200.h declares."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "ppg_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ppg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    from ppg_slam_b200 import capi
    lib = capi.load()
    declared = _header_functions()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), "libppg_b200.so does not export %s" % name
    assert sorted(capi.SYMBOLS) == declared, "capi.SYMBOLS is out of sync with include/ppg_b200.h"
    assert lib.ppg_api_version() == 7


def test_struct_layouts_match_header():
    """ctypes mirrors must have the field order of the header structs."""
    from ppg_slam_b200 import capi
    src = open(os.path.join(ROOT, "include", "ppg_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)

    def fields(struct_name):
        body = re.search(r"typedef struct\s*\{([^{}]*)\}\s*%s;" % struct_name, src, flags=re.S).group(1)
        out = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            names = re.sub(r"\[[^\]]*\]", "", decl).replace("*", " ")
            parts = [p.strip() for p in names.split(",")]
            out.append(parts[0].split()[-1])
            out.extend(p.split()[-1] for p in parts[1:])
        return out

    assert fields("ppg_config") == [f[0] for f in capi.Config._fields_]
    assert fields("ppg_frame_out") == [f[0] for f in capi.FrameOut._fields_]
    assert fields("ppg_assoc_in") == [f[0] for f in capi.AssocIn._fields_]
    assert fields("ppg_assoc_out") == [f[0] for f in capi.AssocOut._fields_]
    assert fields("ppg_map_graph") == [f[0] for f in capi.MapGraph._fields_]
    assert fields("ppg_extend_in") == [f[0] for f in capi.ExtendIn._fields_]
    assert fields("ppg_extend_out") == [f[0] for f in capi.ExtendOut._fields_]
    assert fields("ppg_vocabulary") == [f[0] for f in capi.VocabularyPod._fields_]
    assert fields("ppg_bow_out") == [f[0] for f in capi.BowOut._fields_]
    assert fields("ppg_bow_match_in") == [f[0] for f in capi.BowMatchIn._fields_]
    assert fields("ppg_bow_match_out") == [f[0] for f in capi.BowMatchOut._fields_]
    assert fields("ppg_init_match_in") == [f[0] for f in capi.InitMatchIn._fields_]
    assert fields("ppg_init_match_out") == [f[0] for f in capi.InitMatchOut._fields_]
    assert fields("ppg_triangulation_match_in") == [f[0] for f in capi.TriangulationMatchIn._fields_]
    assert fields("ppg_triangulation_match_out") == [f[0] for f in capi.TriangulationMatchOut._fields_]


def test_no_cpu_fallback():
    """Without a CUDA device the product fails loudly instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from ppg_slam_b200 import cameras, capi
    with pytest.raises(capi.PpgError) as ei:
        capi.Extractor(cameras.EUROC)
    assert ei.value.code == capi.PPG_ERR_CUDA


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "ppg_slam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in text and "from oracle" not in text and "oracle/" not in text.replace(
                    "see oracle", ""), "%s references the oracle" % f
