"""CPU suite: the C-ABI library loads and exports every symbol include/cs_b200.h declares
(no compute calls -- there is no GPU here), and refuses to run without a device."""
import ctypes as C
import os
import re

import pytest

import constraint_solver_b200 as cs
from constraint_solver_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = []
    for fn in sorted(os.listdir(os.path.join(ROOT, "include"))):
        if not fn.endswith(".h"):
            continue
        src = open(os.path.join(ROOT, "include", fn)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names += re.findall(r"\b(cs_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_symbols_all_exported_and_bound():
    lib = cs.load()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
        assert name in L.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(L.SIGNATURES) == declared


def test_abi_version_and_status_strings():
    lib = cs.load()
    assert lib.cs_abi_version() == 4
    assert b"no CPU fallback" in lib.cs_status_string(L.CS_ERR_NO_DEVICE)


def test_host_philox_mirror_matches_oracle():
    from oracle import oracle as orc

    for (seed, chain, purpose, ctr) in [(0, 0, 0, 0), (42, 7, 1, 123456789012), (2**63 + 5, 2**32 - 1, 3, 2**40)]:
        assert cs.philox4x32_10(seed, chain, purpose, ctr) == orc.philox_stream(seed, chain, purpose, ctr)


def test_invalid_config_rejected_before_touching_a_device():
    lib = cs.load()
    h = C.c_void_p()
    bad = L.CsNqConfig(n=0, n_chains=1, chain_offset=0, trace_capacity=0, seed=1, device=-1, neighbourhood=0, flags=0)
    assert lib.cs_nq_create(C.byref(bad), C.byref(h)) == L.CS_ERR_INVALID_ARG
    big = L.CsNqConfig(n=L.CS_NQ_MAX_N + 1, n_chains=1, chain_offset=0, trace_capacity=0, seed=1,
                       device=-1, neighbourhood=0, flags=0)
    assert lib.cs_nq_create(C.byref(big), C.byref(h)) == L.CS_ERR_UNSUPPORTED
    assert lib.cs_nq_create(None, C.byref(h)) == L.CS_ERR_INVALID_ARG
    assert lib.cs_nq_destroy(None) == L.CS_ERR_INVALID_ARG


def test_no_cpu_fallback_without_device():
    if cs.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(cs.CsError) as e:
        cs.NQueensChains(8, 1)
    assert e.value.status == L.CS_ERR_NO_DEVICE


def test_product_package_never_imports_the_oracle():
    """No product source includes, imports, links or loads anything under oracle/ (comments may
    cite it)."""
    pkg = os.path.join(ROOT, "constraint_solver_b200")
    bad = re.compile(r'#\s*include\s*[<"][^>"]*oracle|^\s*(from|import)\s+oracle\b|libcs_oracle|orc_[a-z_]+\s*\(',
                     re.M)
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or fn == "Makefile":
                text = open(os.path.join(dirpath, fn)).read()
                assert not bad.search(text), fn


def test_rust_ffi_declarations_match_the_header():
    """integration/rust/.../ffi.rs cannot be compiled here (no cargo/rustc), so at least keep it in
    lock-step with include/cs_b200.h: every extern fn exists in the header with the same number
    of arguments, and every #[repr(C)] struct lists the header's fields in the header's order."""
    hdr = open(os.path.join(ROOT, "include", "cs_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    rs = open(os.path.join(ROOT, "integration", "rust", "local-search-b200", "src", "ffi.rs")).read()
    rs = re.sub(r"//.*", "", rs)
    c_fns = {m.group(1): [a for a in m.group(2).split(",") if a.strip() and a.strip() != "void"]
             for m in re.finditer(r"\b(cs_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", hdr)}
    rs_fns = {m.group(1): [a for a in m.group(2).split(",") if a.strip()]
              for m in re.finditer(r"pub fn (cs_[a-z0-9_]+)\s*\(([^)]*)\)", rs)}
    assert len(rs_fns) >= 30
    for name, args in rs_fns.items():
        assert name in c_fns, f"ffi.rs declares {name}, which the header does not"
        assert len(args) == len(c_fns[name]), (name, args, c_fns[name])
    c_structs = {m.group(2): re.findall(r"\b([a-z_0-9]+)\s*(?:\[\d+\])?\s*;", m.group(1))
                 for m in re.finditer(r"typedef struct \w+ \{(.*?)\}\s*(\w+)\s*;", hdr, flags=re.S)}
    rs_structs = {m.group(1): re.findall(r"pub ([a-z_0-9]+)\s*:", m.group(2))
                  for m in re.finditer(r"pub struct (cs_\w+)\s*\{(.*?)\}", rs, flags=re.S)}
    checked = 0
    for name, fields in rs_structs.items():
        if name in c_structs:
            assert fields == c_structs[name], (name, fields, c_structs[name])
            checked += 1
    assert checked >= 5
    assert "ABI version %d" % cs.load().cs_abi_version() in open(
        os.path.join(ROOT, "integration", "rust", "local-search-b200", "src", "ffi.rs")).read()


def test_rust_shim_implements_the_reference_trait_surface():
    """the shim implements Solution / Score / SolutionScoreCalculator / MoveProposer /
    InitialSolutionGenerator and offers a LocalSearch-compatible struct with the reference's
    8-argument `new` (local_search.rs:277-299) and `execute` (:301-305).  Source-only (no rustc
    here), so keep its SHAPE under test: the trait impls exist for both problems, `new` lists the
    reference's arguments in the reference's order, and every extern fn it calls is declared."""
    root = os.path.join(ROOT, "integration", "rust", "local-search-b200", "src")
    tr = open(os.path.join(root, "traits.rs")).read()
    ffi = re.sub(r"//.*", "", open(os.path.join(root, "ffi.rs")).read())
    for pat in (r"impl Solution for B200NQueensSolution", r"impl Score for B200NQueensScore",
                r"impl SolutionScoreCalculator for B200NQueensScoreCalculator",
                r"impl<R: rand::Rng> MoveProposer for B200NQueensMoveProposer<R>",
                r"impl<R: rand::Rng> InitialSolutionGenerator for B200NQueensInitialSolutionGenerator<R>",
                r"impl Solution for B200ScheduleSolution", r"impl Score for B200ScheduleScore",
                r"impl SolutionScoreCalculator for B200ScheduleScoreCalculator",
                r"impl<R: rand::Rng> MoveProposer for B200ScheduleMoveProposer<R>",
                r"-> Box<dyn Iterator<Item = Self::Solution>>",
                r"fn get_scored_solution\(&self, solution: Self::_Solution\) -> ScoredSolution<Self::_Solution, Self::_Score>",
                r"pub fn execute\(&mut self, start: _Solution, allow_no_improvement_for: u64\) -> ScoredSolution<_Solution, _Score>"):
        assert re.search(pat, tr), pat
    ref_new = ["move_proposer", "solution_score_calculator", "max_iterations", "window_size",
               "best_solutions_capacity", "all_solutions_capacity", "all_solution_iteration_expiry", "rng"]
    m = re.search(r"impl<R, _Solution, _Score, SSC, MP, DP> B200LocalSearch.*?pub fn new\((.*?)\) -> Self", tr, flags=re.S)
    args = [a.strip().split(":")[0].strip() for a in m.group(1).split(",") if a.strip()]
    assert args[:8] == ref_new, args
    declared = set(re.findall(r"pub fn (cs_[a-z0-9_]+)", ffi))
    used = set(re.findall(r"ffi::(cs_[a-z0-9_]+)\(", tr))
    assert used and used <= declared, used - declared
    assert "pub mod traits;" in open(os.path.join(root, "lib.rs")).read()


def test_bench_helpers_are_deterministic_and_median_is_honest():
    """bench.py host logic that needs no GPU: the synthetic scheduling instances are a pure function
    of (name, seed), and the time-to-best median refuses to summarise runs that mostly failed."""
    import sys

    sys.path.insert(0, ROOT)
    import bench

    w1, ids1, hol1, sk1 = bench.es_instance("es2000x3", 42)
    w2, ids2, hol2, sk2 = bench.es_instance("es2000x3", 42)
    assert (w1, hol1) == (w2, hol2) and list(ids1) == list(ids2) and list(sk1) == list(sk2)
    assert w1["D"] * w1["S"] == 168 and len(hol1) == w1["E"] * w1["nhol"]
    assert all(0 < int(s) < (1 << w1["S"]) for s in sk1)          # everybody qualified for something
    assert bench.es_instance("es50", 42)[3] is None                # reference rotas carry no skill table
    assert bench._median([3.0, 1.0, 2.0]) == 2.0 and bench._median([1.0, None, 3.0, 2.0, 5.0]) == 2.5
    assert bench._median([None, None, 1.0]) is None and bench._median([]) is None
