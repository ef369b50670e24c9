"""Generate tests/golden/reference_signatures.json: the parameter lists (name, default, kind) of every public function
the reference exports from ``nvalchemiops.neighborlist`` (its ``__all__``), read from the reference SOURCES with ``ast``
(the package itself cannot be imported here: ``warp-lang`` is not installed).  Run in the build container:

    python tests/golden/make_reference_signatures.py /root/reference

The fixture is what ``tests/test_oracle_cpu.py::test_public_api_matches_the_reference_signatures`` compares the
drop-in package against; it travels to the GPU box, the reference does not."""
import ast
import json
import os
import sys


def signatures(path):
    out = {}
    for node in ast.parse(open(path).read()).body:
        if not isinstance(node, ast.FunctionDef):
            continue
        a = node.args
        pos = a.posonlyargs + a.args
        defaults = [None] * (len(pos) - len(a.defaults)) + [ast.unparse(d) for d in a.defaults]
        params = [[p.arg, d, "positional_or_keyword"] for p, d in zip(pos, defaults)]
        if a.vararg:
            params.append([a.vararg.arg, None, "var_positional"])
        params += [[p.arg, ast.unparse(d) if d is not None else None, "keyword_only"]
                   for p, d in zip(a.kwonlyargs, a.kw_defaults)]
        if a.kwarg:
            params.append([a.kwarg.arg, None, "var_keyword"])
        out[node.name] = {"params": params, "line": node.lineno}
    return out


def main(ref_root):
    pkg = os.path.join(ref_root, "nvalchemiops", "neighborlist")
    init = ast.parse(open(os.path.join(pkg, "__init__.py")).read())
    exported = {}
    for node in init.body:
        if isinstance(node, ast.ImportFrom) and node.level == 1:
            for alias in node.names:
                exported[alias.name] = node.module + ".py"
    all_names = None
    for node in init.body:
        if isinstance(node, ast.Assign) and any(isinstance(t, ast.Name) and t.id == "__all__" for t in node.targets):
            all_names = [ast.literal_eval(e) for e in node.value.elts]
    result = {"source": "nvalchemiops/neighborlist/__init__.py (__all__) of NVIDIA/nvalchemi-toolkit-ops v0.2.0",
              "functions": {}}
    for name in sorted(all_names):
        mod = exported[name]
        sig = signatures(os.path.join(pkg, mod))[name]
        result["functions"][name] = {"module": mod, "line": sig["line"], "params": sig["params"]}
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_signatures.json")
    with open(dst, "w") as f:
        json.dump(result, f, indent=1)
        f.write("\n")
    print(f"{len(result['functions'])} functions -> {dst}")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
