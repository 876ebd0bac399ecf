"""GPU edge preprocessing (mgcn_preprocess_edges) against the golden vectors written by the UNMODIFIED
reference scripts (data_procs/undirected.py, loop.py, data_add_degree.py via oracle/make_golden.py) and against
the oracle port on random multigraphs: bit-exact edge order, first-occurrence perm, degrees."""
import numpy as np
import pytest
import torch

from meta_gcn_b200 import preprocess as P
from oracle import port
from util import assert_bitexact, golden

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_botnet_pipeline_matches_reference_golden():
    g = golden("preprocess")
    n = int(g["num_nodes"])
    ei, deg = P.botnet_edges(torch.from_numpy(g["raw"]).to(DEV), n)
    assert_bitexact(ei, g["edge_index"], "edge_index")
    assert_bitexact(deg, g["deg"], "deg")


@pytest.mark.parametrize("n,e", [(1, 0), (3, 1), (50, 400), (1000, 30000), (70000, 300000), (5, 4097)])
def test_against_oracle_on_random_multigraphs(n, e):
    rng = np.random.default_rng(n + e)
    raw = rng.integers(0, n, size=(2, e)).astype(np.int64)          # duplicates and self loops included
    d = torch.from_numpy(raw).to(DEV)
    su, perm = P.sort_unique_edges(d, n)
    ref, ref_perm = port.sort_unique_edges(raw, n) if e else (np.zeros((2, 0), np.int64), np.zeros(0, np.int64))
    assert_bitexact(su, ref, "sort_unique_edges")
    assert_bitexact(perm, ref_perm.astype(np.int64), "perm (first occurrence)")
    und, _ = P.to_undirected_ey(d, None, n)
    ref_u = port.to_undirected(raw, n) if e else np.zeros((2, 0), np.int64)
    assert_bitexact(und, ref_u, "to_undirected")
    full, deg = P.botnet_edges(d, n)
    ref_f = port.append_self_loops(ref_u, n)
    assert_bitexact(full, ref_f, "undirected + loops")
    assert_bitexact(deg, port.out_degree(ref_f, n), "deg")
    looped, _ = P.add_self_loops_ey(und, None, None, n)
    assert_bitexact(looped, ref_f, "add_self_loops_ey")


def test_edge_labels_follow_the_permutation():
    rng = np.random.default_rng(7)
    n, e = 200, 3000
    raw = rng.integers(0, n, size=(2, e)).astype(np.int64)
    ey = torch.from_numpy(rng.integers(0, 3, size=e))
    und, uy = P.to_undirected_ey(torch.from_numpy(raw).to(DEV), ey.to(DEV), n)
    both = np.stack([np.concatenate([raw[0], raw[1]]), np.concatenate([raw[1], raw[0]])])
    _, perm = port.sort_unique_edges(both, n)
    assert_bitexact(uy, torch.cat([ey, ey])[torch.from_numpy(perm)], "edge labels")


def test_out_of_range_ids_raise():
    bad = torch.tensor([[0, 5], [1, 2]], dtype=torch.long, device=DEV)
    with pytest.raises(IndexError):
        P.botnet_edges(bad, 3)
