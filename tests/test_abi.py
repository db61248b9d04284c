"""CPU: the C-ABI library loads, exports every symbol include/ccsim.h declares, and struct layouts match."""
import ctypes as C
import importlib
import os
import re
import subprocess

import pytest

abi = importlib.import_module("cluster-capacity_b200._abi")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    hdr = open(os.path.join(ROOT, "include", "ccsim.h")).read()
    return sorted(set(re.findall(r"\b(ccsim_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol(built):
    engine = importlib.import_module("cluster-capacity_b200.engine")
    L = C.CDLL(engine.SO_PATH)
    names = declared_functions()
    assert len(names) >= 12
    for name in names:
        assert hasattr(L, name), name
    assert sorted(engine.EXPORTS) == names
    assert L.ccsim_abi_version() == abi.ABI_VERSION


def test_host_library_exports_every_declared_symbol(built):
    """include/cchost.h: every cc_* entry point the header declares is exported by libcchost.so."""
    hdr = open(os.path.join(ROOT, "include", "cchost.h")).read()
    names = sorted(set(re.findall(r"\b(cc_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 8, names
    L = C.CDLL(os.path.join(ROOT, "cluster-capacity_b200", "libcchost.so"))
    for name in names:
        assert hasattr(L, name), name


def test_struct_layouts_match_header(built, tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "%s/include/ccsim.h"\n'
                   'int main(){printf("%%zu %%zu %%zu %%zu %%zu %%zu %%zu %%d\\n",sizeof(ccsim_template),sizeof(ccsim_nodes),'
                   'sizeof(ccsim_result),sizeof(ccsim_counter),sizeof(ccsim_config),offsetof(ccsim_template,pts),'
                   'offsetof(ccsim_result,pod_node),CCSIM_R_TOTAL);return 0;}\n' % ROOT)
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split()
    want = [C.sizeof(abi.Template), C.sizeof(abi.Nodes), C.sizeof(abi.Result), C.sizeof(abi.Counter), C.sizeof(abi.Config),
            abi.Template.pts.offset, abi.Result.pod_node.offset, abi.R_TOTAL]
    assert [int(x) for x in out] == want


def test_no_cpu_fallback_without_gpu(built):
    """On a host without a CUDA device the product path must fail loudly, never fall back to the oracle."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    engine = importlib.import_module("cluster-capacity_b200.engine")
    with pytest.raises(engine.EngineError):
        engine.Engine(device=0)


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "cluster-capacity_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "from oracle" not in text and "import oracle" not in text and "ccsim_oracle" not in text, f
