"""Regenerates tests/golden/ from the reference checkout (run in the build container only;
/root/reference does not exist on the GPU box, which is why the outputs are committed).

1. Copies the reference's own least-squares test fixtures (data files, not source):
     src/test/resources/{aMat,bMat,aMatShuffled,bMatShuffled,aMat-1class,bMat-1class}.csv
   used by T/nodes/learning/BlockWeightedLeastSquaresSuite.scala.
2. Writes golden.json: constants quoted from the reference test suites (file:line in each
   entry) plus oracle outputs on the fixtures, so later oracle edits are caught as diffs.
"""
import json, os, shutil, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/src/test/resources"
sys.path.insert(0, ROOT)
from oracle import keystone_oracle as ko  # noqa: E402

FIXTURES = ["aMat.csv", "bMat.csv", "aMatShuffled.csv", "bMatShuffled.csv", "aMat-1class.csv", "bMat-1class.csv"]


def convolver_fixture():
    """3. conv_gantrycrane.npz: the reference's convolution test image (images/gantrycrane.png, decoded to 8-bit RGB) and channel 0
    of the expected convolution (images/convolved.gantrycrane.csv: "x,y,value" lines of a ColumnMajorArrayVectorizedImage) --
    T/nodes/images/ConvolverSuite.scala:100-137 ("convolutions should match scipy")."""
    from PIL import Image
    rgb = np.array(Image.open(os.path.join(REF, "images", "gantrycrane.png")).convert("RGB"), dtype=np.uint8)
    raw = np.loadtxt(os.path.join(REF, "images", "convolved.gantrycrane.csv"), delimiter=",")
    xd, yd = int(raw[:, 0].max()) + 1, int(raw[:, 1].max()) + 1
    expected = raw[:, 2].reshape(xd, yd)          # value (x, y) at y + x * yDim
    assert np.array_equal(expected, np.rint(expected))
    np.savez_compressed(os.path.join(HERE, "conv_gantrycrane.npz"), rgb=rgb, expected=expected.astype(np.int32))


def main():
    convolver_fixture()
    for f in FIXTURES:
        shutil.copyfile(os.path.join(REF, f), os.path.join(HERE, f))
    A = np.loadtxt(os.path.join(HERE, "aMat.csv"), delimiter=",")
    B = np.loadtxt(os.path.join(HERE, "bMat.csv"), delimiter=",")
    out = {
        "standard_scaler": {
            "cite": "T/nodes/stats/StandardScalerSuite.scala:21-28,58-59",
            "dense_data": [[-2.0, 2.3, 0.0], [0.0, -1.0, -3.0], [0.0, -5.1, 0.0], [3.8, 0.0, 1.9], [1.7, -0.6, 0.0], [0.0, 1.9, 0.0]],
            "row0": [-1.31527964, 1.023470449, 0.11637768424],
            "row3": [1.637735298, 0.156973995, 1.32247368462],
            "tol": 1e-5,
        },
        "linear_mapper": {
            "cite": "T/nodes/learning/LinearMapperSuite.scala:13-36",
            "x": [5.0, 4.0, 3.0, 2.0, -1.0], "point": [2.0, -3.0, 2.0, 3.0, 5.0], "expected": 5.0, "tol": 1e-8,
        },
        "bwls": {
            "cite": "T/nodes/learning/BlockWeightedLeastSquaresSuite.scala:115-223",
            "lambda": 0.1, "mixture_weight": 0.3,
            "grad_tol_b4": 1e-2, "grad_tol_b5": 1e-1, "pcs_match_tol": 1e-6,
        },
    }
    # oracle outputs on the fixtures (regression pins for the oracle itself)
    for b, iters in ((4, 10), (5, 10), (4, 5)):
        xs, fb = ko.bwls_fit(A, B, b, iters, 0.1, 0.3)
        W = np.concatenate(xs, 0)
        g = ko.compute_gradient(A, B, 0.1, 0.3, W, fb)
        out["bwls"][f"b{b}_it{iters}"] = {"W": W.tolist(), "final_b": fb.tolist(), "grad_norm": float(np.linalg.norm(g))}
    xs, yb, mus = ko.block_ls_fit(A, B, 4, 3, 0.1)
    out["block_ls_fixture"] = {"note": "oracle output, mlmatrix boundary unpinned", "block_size": 4, "num_iter": 3,
                               "lambda": 0.1, "W": np.concatenate(xs, 0).tolist(), "intercept": yb.tolist()}
    with open(os.path.join(HERE, "golden.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", os.path.join(HERE, "golden.json"))


if __name__ == "__main__":
    main()
