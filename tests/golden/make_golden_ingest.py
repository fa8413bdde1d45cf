"""Generate tests/golden/ingest.npz by EVALUATING the reference's own normalisation expression
(read from /root/reference at generation time by `ast`, never copied):

    cub/code/data/data.py:133-135  (DataAugmentation.stochastic_appearance_augmentation)
        output_images = [o.astype(np.float32) * 2.0 / 255.0 - 1.0 for o in output_images]

Run in the build container only:  python tests/golden/make_golden_ingest.py
"""
import ast
import os

import numpy as np

REF = "/root/reference/cub/code/data/data.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def reference_expression():
    tree = ast.parse(open(REF).read())
    for fn in ast.walk(tree):
        if isinstance(fn, ast.FunctionDef) and fn.name == "stochastic_appearance_augmentation":
            for node in ast.walk(fn):
                if isinstance(node, ast.ListComp) and "astype(np.float32)" in ast.unparse(node.elt):
                    assert node.generators[0].target.id == "o"
                    return ast.unparse(node.elt), compile(ast.Expression(node.elt), REF, "eval")
    raise RuntimeError("normalisation expression not found")


def main():
    text, code = reference_expression()
    print("reference expression:", text)
    all_bytes = np.arange(256, dtype=np.uint8)
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, size=(2, 3, 16, 16, 3), dtype=np.uint8)   # [V,B,S,S,3]
    f = lambda o: eval(code, dict(np=np, o=o))                            # noqa: E731
    out_bytes, out_img = f(all_bytes), f(img)
    assert out_bytes.dtype == np.float32 and out_img.dtype == np.float32
    np.savez_compressed(os.path.join(HERE, "ingest.npz"), expression=np.array(text), all_bytes=all_bytes,
                        out_bytes=out_bytes, img=img, out_img=out_img)


if __name__ == "__main__":
    main()
