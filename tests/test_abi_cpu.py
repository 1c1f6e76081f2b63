"""CPU: the C-ABI library builds, loads, and exports exactly what include/lgcn_b200.h declares;
ctypes mirrors of the structs have the C layout; host-only entry points work; device entry points
refuse CPU tensors (no fallback)."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

import lgcn_b200  # noqa: F401
from lgcn_b200 import _lib
from lgcn_b200.data import dataset_handler as dh
from lgcn_b200.data import synthetic
from oracle import pyg_restated as pyg

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(REPO, "include", "lgcn_b200.h")


@pytest.fixture(scope="module")
def L():
    import __graft_entry__ as ge
    ge.build()
    return _lib.lib()


def test_header_and_library_agree(L):
    src = open(HEADER).read()
    declared = set(re.findall(r"^\s*(?:const char \*|int|size_t)\s*\*?\s*(lgcn_\w+)\s*\(", src, flags=re.M))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.lgcn_version() == 200
    assert isinstance(L.lgcn_last_error(), bytes)


def test_struct_layouts_match_c(tmp_path, L):
    prog = tmp_path / "sz.c"
    prog.write_text('#include <stdio.h>\n#include "lgcn_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                    'sizeof(lgcn_task),sizeof(lgcn_graph),sizeof(lgcn_graph_sizes),sizeof(lgcn_adam),'
                    'sizeof(lgcn_step_buffers),offsetof(lgcn_graph,partials),sizeof(lgcn_bpr_owner_ws),'
                    'sizeof(lgcn_peers),offsetof(lgcn_bpr_owner_ws,scalars));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(REPO, "include"), str(prog), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    got = [ctypes.sizeof(_lib.CTask), ctypes.sizeof(_lib.CGraph), ctypes.sizeof(_lib.CGraphSizes),
           ctypes.sizeof(_lib.CAdam), ctypes.sizeof(_lib.CStepBuffers), _lib.CGraph.partials.offset,
           ctypes.sizeof(_lib.CBprOwnerWs), ctypes.sizeof(_lib.CPeers), _lib.CBprOwnerWs.scalars.offset]
    assert [int(x) for x in out] == got


def test_sm100a_cubin_present():
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_host_metis_entry_point_matches_oracle_call(L):
    g = synthetic.make_graph("tiny", seed=4)
    train = g.edges("train")
    part = dh.metis_partition(train, g.num_nodes, 6)
    se, _ = pyg.sort_edge_index(train, g.num_nodes)
    want = pyg.metis_partition(pyg.index2ptr(se[0], g.num_nodes), se[1], 6)
    assert torch.equal(part, want) and int(part.max()) == 5
    assert L.lgcn_partition_metis(0, None, None, 4, None) == -1           # LGCN_E_INVALID


def test_device_entry_points_refuse_cpu_tensors(L):
    from lgcn_b200.models.light_gcn import LightGCN
    from lgcn_b200.utils.recommend import score_topk
    m = LightGCN(5, 7, num_layers=2)
    with pytest.raises(_lib.LgcnError):
        m(torch.tensor([[0], [5]]))
    with pytest.raises(_lib.LgcnError):
        score_topk(torch.zeros(2, 64), torch.zeros(3, 64), 2)
    with pytest.raises(_lib.LgcnError):
        LightGCN(5, 7, dim_h=32)


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/liblgcn_b200.so")
    with pytest.raises(_lib.LgcnError, match="no CPU or PyTorch fallback"):
        _lib.lib()
