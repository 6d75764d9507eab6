"""Golden vectors for the anchor decode (reference: det3d/core/bbox/box_torch_ops.py:79-148, called through
GroundBox3dCoderTorch.decode_torch, box_coders.py:106-109).  Run in the development container (needs /root/reference):

    python tests/golden/make_golden_decode.py

The reference function is lifted from its source file with `ast` (the module imports compiled extensions that are not
installed) and executed as is; nothing of it is stored here.  The coder passes `linear_dim` POSITIONALLY into the ignored
`bin_loss` slot and never forwards `norm_velo`: the golden is generated through exactly that call."""
import ast
import os

import numpy as np
import torch

REF = "/root/reference/det3d/core/bbox/box_torch_ops.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "decode_coder.npz")


def lift(path, name):
    src = open(path).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == name)
    ns = {"torch": torch, "np": np}
    exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return ns[name]


def main():
    decode = lift(REF, "second_box_decode")
    rng = np.random.default_rng(7)
    out = {}
    for nd in (7, 9):
        anchors = rng.normal(size=(257, nd)).astype(np.float32)
        anchors[:, 3:6] = rng.uniform(0.5, 4.0, size=(257, 3)).astype(np.float32)      # sizes > 0
        for vec in (False, True):
            enc = (rng.normal(size=(257, nd + (1 if vec else 0))) * 0.3).astype(np.float32)
            for linear_dim in (False, True):
                # the reference coder's call: decode(boxes, anchors, self.vec_encode, self.linear_dim)
                got = decode(torch.from_numpy(enc), torch.from_numpy(anchors), vec, linear_dim).numpy()
                key = "nd%d_vec%d_lin%d" % (nd, int(vec), int(linear_dim))
                out[key + "_enc"], out[key + "_anchors"], out[key + "_out"] = enc, anchors, got
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, len(out) // 3, "cases")


if __name__ == "__main__":
    main()
