"""Static checks of bench.py that run without a GPU: every `"..." % args` expression has a well-formed format string
whose number of conversions matches the argument tuple (a stray `%` in a note string once broke the JSON line on the
GPU box only, after all measurements had run), and the reference arm runs on the CPU."""
import ast
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SPEC = re.compile(r"%(?:\([^)]*\))?[#0\- +]*(?:\*|\d+)?(?:\.(?:\*|\d+))?[hlL]?([diouxXeEfFgGcrsa%])")


def _joined(node):
    """the constant string of a (possibly implicitly concatenated / '+'-joined) literal, else None"""
    if isinstance(node, ast.Constant) and isinstance(node.value, str):
        return node.value
    if isinstance(node, ast.BinOp) and isinstance(node.op, ast.Add):
        a, b = _joined(node.left), _joined(node.right)
        return a + b if a is not None and b is not None else None
    return None


def test_every_percent_format_in_bench_is_well_formed():
    for fn in ("bench.py", "__graft_entry__.py"):
        tree = ast.parse(open(os.path.join(ROOT, fn)).read())
        checked = 0
        for node in ast.walk(tree):
            if not (isinstance(node, ast.BinOp) and isinstance(node.op, ast.Mod)):
                continue
            fmt = _joined(node.left)
            if fmt is None:
                continue
            # every '%' must start a valid conversion
            rest = SPEC.sub("", fmt)
            assert "%" not in rest, (fn, node.lineno, fmt[:80])
            n = sum(1 for m in SPEC.finditer(fmt) if m.group(1) != "%")
            if isinstance(node.right, ast.Tuple):
                assert n == len(node.right.elts), (fn, node.lineno, n, len(node.right.elts), fmt[:80])
            elif not isinstance(node.right, (ast.Dict, ast.Name, ast.Call, ast.Attribute, ast.Subscript, ast.BinOp, ast.IfExp)):
                assert n == 1, (fn, node.lineno, fmt[:80])
            checked += 1
        assert checked > 0 or fn != "bench.py"


def test_reference_arm_prints_one_json_line_on_cpu():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--cpu-seconds", "0.3"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["higher_is_better"] is True
