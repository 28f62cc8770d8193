"""Generates tests/golden/preproc_depth.npz from the UNMODIFIED ``Agent_Helper._preprocess_depth`` of
/root/reference/nav/agent/agent_helper.py (extracted by ast - the module itself needs skimage/habitat, absent here),
next to oracle/preproc.py on the same seeded inputs; refuses to write unless the two agree bit for bit.
Run in the build container only:  python tests/golden/make_preproc_golden.py"""
import ast
import os
import sys
import textwrap

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import preproc as oracle  # noqa: E402

SRC = "/root/reference/nav/agent/agent_helper.py"


def reference_function(name):
    text = open(SRC).read()
    tree = ast.parse(text)
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == name:
            src = textwrap.dedent(ast.get_source_segment(text, node))
            ns = {"np": np}
            exec(compile(src, SRC, "exec"), ns)
            return ns[name]
    raise KeyError(name)


def main():
    ref = reference_function("_preprocess_depth")
    seeds = [0, 1, 2]
    outs = []
    for s in seeds:
        d = oracle.synth_depth(s)
        r = ref(None, d.copy(), 0.5, 5.0)
        o = oracle.preprocess_depth(d.copy(), 0.5, 5.0)
        assert r.dtype == o.dtype == np.float32 and np.array_equal(r, o), f"seed {s}: oracle differs from the reference"
        outs.append(r[2::4, 2::4].astype(np.float32))
    np.savez_compressed(os.path.join(HERE, "preproc_depth.npz"), seeds=np.array(seeds), depth_cm_sub=np.stack(outs),
                        full_sum=np.array([float(ref(None, oracle.synth_depth(s), 0.5, 5.0).astype(np.float64).sum()) for s in seeds]))
    print("ok", [int((o == 45050.0).sum()) for o in outs])


if __name__ == "__main__":
    main()
