"""C++ host mirror (include/local_search_b200.hpp) and the C++ drivers (examples/cpp/).

The reference is Rust and there is no Rust toolchain here, so the host side above the C ABI is
C++; tests/cpp/host_mirror_test.cpp holds the assertions (it reads like the reference's own
tests, examples/nqueens/src/lib.rs:94-119 and main.rs:157-200) and links the CPU oracle as the
checker.  This file compiles it, runs the host-only part on CPU and the parity part on the GPU.
"""
import hashlib
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "cpp", "build")
LIBDIR = os.path.join(ROOT, "constraint_solver_b200")
ORCDIR = os.path.join(ROOT, "oracle")


def _compile(src, out, with_oracle=False):
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, out)
    deps = [src, os.path.join(ROOT, "include", "local_search_b200.hpp"), os.path.join(ROOT, "include", "cs_b200.h"),
            os.path.join(ROOT, "examples", "cpp", "solver_context.hpp")]
    if os.path.exists(exe) and all(os.path.getmtime(exe) >= os.path.getmtime(d) for d in deps):
        return exe
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"), src,
           "-L" + LIBDIR, "-lcs_b200", "-Wl,-rpath," + LIBDIR, "-o", exe]
    if with_oracle:
        cmd += ["-I" + ORCDIR, "-L" + ORCDIR, "-lcs_oracle", "-Wl,-rpath," + ORCDIR]
    subprocess.check_call(cmd)
    return exe


@pytest.fixture(scope="module")
def host_test_exe():
    return _compile(os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp"), "host_mirror_test", with_oracle=True)


@pytest.fixture(scope="module")
def nqueens_exe():
    return _compile(os.path.join(ROOT, "examples", "cpp", "nqueens_main.cpp"), "nqueens_b200")


@pytest.fixture(scope="module")
def scheduling_exe():
    return _compile(os.path.join(ROOT, "examples", "cpp", "employee_scheduling_main.cpp"), "employee_scheduling_b200")


def test_cpp_host_cpu(host_test_exe):
    out = subprocess.run([host_test_exe, "--cpu"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    # hash_str == BLAKE2b-256 (examples/nqueens/src/main.rs:28-33) for every printed message
    msgs = {0: b"", 3: b"abc"}
    seen = 0
    for line in out.stdout.splitlines():
        if line.startswith("blake2b256 "):
            _, ln, hx = line.split()
            seen += 1
            if int(ln) in msgs:
                assert hx == hashlib.blake2b(msgs[int(ln)], digest_size=32).hexdigest()
    assert seen == 5
    lines = [l.split()[2] for l in out.stdout.splitlines() if l.startswith("blake2b256 2 ")]
    assert lines == [hashlib.blake2b(s, digest_size=32).hexdigest() for s in (b"42", b"43")]
    long = (b"the quick brown fox jumps over the lazy dog, " * 2 +
            b"the quick brown fox jumps over the lazy dog -- longer than one 128-byte block")
    assert len(long) == 167
    assert f"blake2b256 167 {hashlib.blake2b(long, digest_size=32).hexdigest()}" in out.stdout


def test_cpp_drivers_build_and_fail_loudly_without_gpu(nqueens_exe, scheduling_exe):
    import constraint_solver_b200 as cs

    if cs.device_count() > 0:
        pytest.skip("a GPU is present; the loud-failure path is for GPU-less hosts")
    for exe in (nqueens_exe, scheduling_exe):
        out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
        assert out.returncode == 101 and "no CUDA device" in out.stderr
    out = subprocess.run([nqueens_exe, "--board-size", "x8"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 2  # clap's validator rejects a non-integer board size (main.rs:113-118)


@pytest.mark.gpu
def test_cpp_host_gpu(host_test_exe):
    out = subprocess.run([host_test_exe, "--gpu"], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "checks passed" in out.stdout


@pytest.mark.gpu
def test_cpp_nqueens_driver(nqueens_exe):
    """examples/nqueens/src/main.rs: default run (seed "42", board 8) ends on score 0."""
    out = subprocess.run([nqueens_exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    assert out.stdout.startswith("local search n-queens example\n")
    assert "result.score: NQueensScore(0)" in out.stdout
    board = out.stdout.split("result.solution:\n")[1].split("\nresult.score")[0]
    assert board.count("Q") == 8 and len(board.splitlines()) == 17
    out2 = subprocess.run([nqueens_exe, "-s", "42", "-b", "8"], capture_output=True, text=True, timeout=600)
    assert out2.stdout == out.stdout  # run-to-run determinism (`repeatable`)
    big = subprocess.run([nqueens_exe, "--board-size", "64", "--chains", "256"], capture_output=True, text=True, timeout=600)
    assert big.returncode == 0 and "NQueensScore(0)" in big.stdout


@pytest.mark.gpu
def test_cpp_scheduling_driver_json(scheduling_exe, tmp_path):
    """wasm JSON shapes (web/employee-scheduling-wasm-bindgen/src/lib.rs:86-110) end to end."""
    import datetime as dt

    from oracle import oracle as orc

    inp = {"startDate": "2022-05-09", "endDate": "2022-06-05", "employees": [{"id": i} for i in (3, 1, 4, 15, 9, 2, 6)],
           "employeeHolidays": [["2022-05-10", "2022-05-11"], [], ["2022-06-01"], [], [], [], []]}
    p = tmp_path / "in.json"
    p.write_text(json.dumps(inp))
    out = subprocess.run([scheduling_exe, "--json", str(p), "--chains", "64", "--progress"], capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stderr
    info = [json.loads(l) for l in out.stderr.splitlines() if l.startswith("{")]   # get_iteration_info per round
    assert info and info[0] == {"current": 1, "total": 250} and [i["current"] for i in info] == list(range(1, len(info) + 1))
    res = json.loads(out.stdout)
    days = res["days_to_employees"]
    assert len(days) == 28 and days[0][0] == "Mon 2022-05-09" and days[-1][0] == "Sun 2022-06-05"
    ids = [d[1]["id"] for d in days]
    assert set(ids) <= {3, 1, 4, 15, 9, 2, 6}
    start = dt.date(2022, 5, 9)
    hol = [(e["id"], (dt.date.fromisoformat(h) - start).days) for e, hs in zip(inp["employees"], inp["employeeHolidays"]) for h in hs]
    hard, soft = orc.es_score(ids, 0, hol)
    assert res["score"] == {"hard_score": float(hard), "soft_score": float(soft)}
    assert hard == 0
    plain = subprocess.run([scheduling_exe], capture_output=True, text=True, timeout=600)
    assert plain.returncode == 0 and "result.score: ScheduleScore { hard_score: OrderedFloat(0.0)" in plain.stdout
    assert plain.stdout.count("employee: Employee { id:") >= 2


@pytest.mark.gpu
def test_cpp_multi_gpu_driver_single_process_nccl():
    """examples/cpp/nqueens_multi_gpu.cpp: one process, one handle per GPU, NCCL on the library's
    device pointers.  The partitioned scan must give the same trajectory on 1 and on all GPUs."""
    import re
    import shutil

    import constraint_solver_b200 as cs

    if shutil.which("make") is None or not os.path.exists("/usr/include/nccl.h"):
        pytest.skip("needs make and nccl.h")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "examples", "cpp"), "nqueens_multi_gpu"])
    exe = os.path.join(ROOT, "examples", "cpp", "nqueens_multi_gpu")
    G = cs.device_count()

    def run(*args):
        out = subprocess.run([exe, *args], capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stdout + out.stderr
        m = re.search(r"(\S+) n=(\d+) gpus=(\d+) steps=(\d+): (\S+) moves scored .* best score (-?\d+)", out.stdout)
        assert m, out.stdout
        return int(m.group(5)), int(m.group(6))

    n, steps = 3000, 6
    moves1, best1 = run("--partitioned", "--board-size", str(n), "--steps", str(steps), "--gpus", "1")
    assert moves1 == steps * n * (n - 1) // 2
    if G >= 2:
        movesg, bestg = run("--partitioned", "--board-size", str(n), "--steps", str(steps), "--gpus", str(G))
        assert (movesg, bestg) == (moves1, best1)
    moves, best = run("--board-size", "500", "--chains", "64", "--steps", "4", "--exchange", "2", "--gpus", str(G))
    assert moves == 4 * G * 64 * (500 * 499 // 2) and best >= 0
