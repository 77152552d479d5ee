"""CPU-side checks of the C-ABI library and the host layer: every symbol include/ksched.h declares is exported,
struct sizes match the header, the encoder produces the documented shapes, unsupported inputs fail LOUDLY, and without
a GPU the solver refuses to run (there is no CPU fallback in the product)."""
import ctypes as C
import re
from pathlib import Path

import pytest

import fixtures as fx

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_header_symbol(pkg):
    header = (ROOT / "include" / "ksched.h").read_text()
    declared = set(re.findall(r"^(?:int|void|const char\*)\s+(ksched_\w+)\(", header, re.M))
    assert declared == set(pkg.ABI_SYMBOLS), declared ^ set(pkg.ABI_SYMBOLS)
    lib = pkg.lib()
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert lib.ksched_abi_version() == int(re.search(r"#define\s+KSCHED_ABI_VERSION\s+(\d+)", header).group(1)) == 4
    assert lib.ksched_type_words(1000) == 16 and lib.ksched_type_words(1) == 1 and lib.ksched_type_words(65) == 2


def test_no_torch_types_in_the_abi():
    header = (ROOT / "include" / "ksched.h").read_text()
    assert "torch" not in header and "at::" not in header and "std::" not in header


def test_shard_ranges_partition_the_columns(pkg):
    for words in (1, 7, 32, 33, 250):
        for world in (1, 2, 4, 8):
            spans = [pkg.shard_range(words, r, world) for r in range(world)]
            covered = [w for b, e in spans for w in range(b, e)]
            assert covered == list(range(words)), (words, world, spans)


def test_quantity_parser(pkg):
    q = pkg.model_lib().kh_parse_quantity
    assert q(b"100m") == 100 and q(b"1") == 1000 and q(b"1.8G") == 1_800_000_000_000 and q(b"4Gi") == 4 * 1024 ** 3 * 1000
    assert q(b"10Mi") == 10 * 1024 ** 2 * 1000 and q(b"2Ti") == 2 * 1024 ** 4 * 1000 and q(b"4.5") == 4500 and q(b"1m") == 1
    assert q(b"1u") == -2 ** 63  # sub-milli quantities are rejected, not rounded


@pytest.mark.parametrize("config,pods,types,nodes", [(1, 100, 10, 0), (2, 500, 500, 0), (3, 500, 1000, 0), (4, 500, 1000, 0), (5, 200, 1000, 20)])
def test_encoder_shapes(pkg, config, pods, types, nodes):
    problem = pkg.Problem.synth(config, pods, types, 42, nodes)
    rs = pkg.ResidentSolve(problem)
    d = rs.dims
    assert d["types"] == types and d["type_words"] == (types + 63) // 64
    assert d["pods"] == (0 if config == 5 else pods)
    assert d["existing"] == nodes
    assert d["templates"] == (3 if config == 3 else 1)
    assert d["keys"] <= 16 and d["resources"] == 3
    if config == 2:
        assert d["classes"] <= 30  # 5 cpu x 6 memory request classes
    if config == 4:
        assert d["groups"] > 0


def test_library_exports_every_host_header_symbol(pkg):
    """include/ksched_host.h: the string-level entry points a cgo shim binds when it lets the library encode"""
    header = (ROOT / "include" / "ksched_host.h").read_text()
    declared = set(re.findall(r"\b(kh_[a-z0-9_]+)\(", header))
    assert len(declared) >= 25
    lib, model = pkg.lib(), pkg.model_lib()   # link with -lksched -lkmodel
    for sym in declared:
        assert hasattr(lib, sym) or hasattr(model, sym), sym
    import subprocess
    src = "#include \"ksched_host.h\"\nint main(void) { return 0; }\n"
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", str(ROOT / "include"), "-x", "c", "-"], input=src, text=True, capture_output=True)
    assert r.returncode == 0, r.stderr     # both headers are plain C


def test_duplicate_uids_are_rejected(pkg, oracle):
    prob = fx.problem([fx.pod({"cpu": "1"}, uid="same"), fx.pod({"cpu": "1"}, uid="same")])
    p = pkg.Problem.from_dict(prob)
    with pytest.raises(pkg.KschedError):
        pkg.ResidentSolve(p)
    res = pkg.Result()
    assert oracle.solve(p, res) != 0 and "unique UIDs" in res.error


def test_unsupported_inputs_fail_loudly(pkg):
    # what the encoding cannot express is refused, never approximated: more than 16 active label keys, more than 7 distinct
    # Gt/Lt thresholds on one key, a key whose values + regions do not fit one 64-bit word
    many = fx.pod({"cpu": "1"}, nodeSelector={f"custom-{i}": "x" for i in range(17)})
    with pytest.raises(pkg.KschedError) as e:
        pkg.ResidentSolve(pkg.Problem.from_dict(fx.problem([many])))
    assert "unsupported" in str(e.value)
    thr = [fx.pod({"cpu": "1"}, nodeAffinity={"required": [[{"key": "integer", "operator": "Gt", "values": [str(i)]}]]}) for i in range(8)]
    with pytest.raises(pkg.KschedError) as e:
        pkg.ResidentSolve(pkg.Problem.from_dict(fx.problem(thr)))
    assert "unsupported" in str(e.value)
    wide = fx.pod({"cpu": "1"}, nodeAffinity={"required": [[{"key": "team", "operator": "In", "values": [f"v{i}" for i in range(64)]}]]})
    with pytest.raises(pkg.KschedError) as e:
        pkg.ResidentSolve(pkg.Problem.from_dict(fx.problem([wide])))
    assert "unsupported" in str(e.value)


def test_gt_lt_and_complement_instance_types_are_encoded(pkg):
    """the encoder hands Gt/Lt over in region form (include/ksched.h: ksched_key_regions) instead of refusing them"""
    its = fx.default_instance_types()
    its[0]["requirements"].append({"key": "custom", "operator": "NotIn", "values": ["x"]})
    pr = fx.provisioner(labels={"custom": "y"}, requirements=[{"key": "integer", "operator": "Gt", "values": ["1"]}])
    gt = fx.pod({"cpu": "1"}, nodeAffinity={"required": [[{"key": "integer", "operator": "Lt", "values": ["20"]}]]})
    rs = pkg.ResidentSolve(pkg.Problem.from_dict(fx.problem([gt], instance_types=its, provisioners=[pr])))
    assert rs.dims["pods"] == 1


def test_no_cpu_fallback(pkg):
    """On a box without a CUDA device the product must refuse, not silently compute on the CPU."""
    if pkg.device_count() > 0:
        pytest.skip("a GPU is present")
    problem = pkg.Problem.synth(1, 100, 10, 42, 0)
    with pytest.raises(pkg.KschedError) as e:
        pkg.Scheduler(problem).solve()
    assert e.value.code == pkg.KSCHED_ERR_NO_DEVICE


def test_model_library_has_no_solver_and_no_cuda():
    """libkmodel.so (what the oracle's tests and the bench's reference arm load) is the string-level model only"""
    import subprocess
    so = ROOT / "karpenter-core_b200" / "libkmodel.so"
    out = subprocess.run(["ldd", str(so)], capture_output=True, text=True).stdout
    assert "cuda" not in out.lower() and "nccl" not in out.lower() and "ksched" not in out
    syms = subprocess.run(["nm", "-D", "--defined-only", str(so)], capture_output=True, text=True).stdout
    assert "ksched_solve" not in syms and "kh_scheduler_solve" not in syms and "kh_problem_from_json" in syms


def test_product_does_not_link_or_import_the_oracle():
    import subprocess
    out = subprocess.run(["ldd", str(ROOT / "karpenter-core_b200" / "libksched.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out
    exts = ("*.cc", "*.cu", "*.cuh", "*.h", "*.py")
    for path in [p for e in exts for p in (ROOT / "karpenter-core_b200").rglob(e)]:
        text = path.read_text()
        for needle in ('#include "../oracle', "#include \"oracle", "liboracle", "oracle_lib", "import oracle", "oracle_solve"):
            assert needle not in text, (path, needle)
