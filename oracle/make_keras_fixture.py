"""Pin the Keras half of the oracle: run the UNMODIFIED reference graphs under TensorFlow and store their outputs.

TEST INFRASTRUCTURE ONLY.  TensorFlow / Keras cannot be installed in the build image (no network), so this script is
meant for any machine that has `tensorflow==2.11` (requirements.txt:4 of the reference) and a checkout of
WenChentao/3DeeCellTracker:

    python -m oracle.make_keras_fixture --reference /path/to/3DeeCellTracker [--out tests/golden/keras_fixture.npz]

It imports CellTracker.unet3d / ffn / preprocess as they are, sets the seeded weights of oracle/unet.py and
oracle/ffn.py (Keras `get_weights()` order), runs seeded inputs and writes inputs + outputs.  When the file exists,
tests/test_oracle_golden.py::test_keras_fixture_* compare the torch restatement with it (CPU) and
tests/test_gpu_lcn_unet.py::test_gpu_matches_keras_fixture compares the CUDA path with it; both skip loudly when it
is absent.  Until a fixture is committed the Keras half stays "parity unpinned" (DESIGN.md section 4).

The same machine can convert the authors' pretrained models for the product:
    python -m 3deecelltracker_b200.io_formats unet3_pretrained.h5 unet3_pretrained.npz        (needs only h5py)
"""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", required=True, help="checkout of WenChentao/3DeeCellTracker")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "keras_fixture.npz"))
    args = ap.parse_args()
    sys.path.insert(0, args.reference)
    import tensorflow as tf                                   # noqa: F401  (fails loudly where TF is missing)
    from CellTracker import ffn as ref_ffn
    from CellTracker import preprocess as ref_pre
    from CellTracker import unet3d as ref_unet
    from oracle import ffn as offn
    from oracle import unet as ounet

    out = {"tf_version": np.array(tf.__version__)}
    rng = np.random.default_rng(2024)

    # ---- unet3_a / b / c (unet3d.py:26-81): one tile each, seeded weights in get_weights() order
    for variant, builder in (("a", ref_unet.unet3_a), ("b", ref_unet.unet3_b), ("c", ref_unet.unet3_c)):
        model = builder()
        ws = ounet.random_weights(variant, seed=7)
        assert [w.shape for w in model.get_weights()] == [w.shape for w in ws], "weight order differs from the oracle's"
        model.set_weights(ws)
        shape = tuple(model.input_shape[1:4])
        x = rng.normal(0, 1, (1,) + shape + (1,)).astype(np.float32)
        out[f"unet_{variant}__x"] = x
        out[f"unet_{variant}__y"] = np.asarray(model.predict(x), dtype=np.float32)
        if variant == "a":
            vol = rng.normal(0, 1, (1, 100, 90, 20, 1)).astype(np.float32)
            out["prediction_a__img"] = vol
            out["prediction_a__out"] = np.asarray(ref_unet.unet3_prediction(vol, model, shrink=(24, 24, 2)), dtype=np.float32)

    # ---- FFN (ffn.py:225-265): dense (B,122) form
    ffn_model = ref_ffn.FFN()
    x = rng.normal(0, 0.7, (256, 122)).astype(np.float32)
    ffn_model(x)                                              # builds the variables
    ws = offn.random_weights(7)
    assert [w.shape for w in ffn_model.get_weights()] == [w.shape for w in ws], "FFN weight order differs"
    ffn_model.set_weights(ws)
    out["ffn__x"] = x
    out["ffn__y"] = np.asarray(ffn_model.predict(x, batch_size=1024), dtype=np.float32)

    # ---- LCN (preprocess.py:117-188)
    raw = np.clip(rng.normal(100, 30, (64, 64, 16)), 0, 65535).astype(np.uint16)
    out["lcn__raw"] = raw
    out["lcn__normalize_image"] = np.asarray(ref_pre._normalize_image(raw.copy(), 20))
    img = np.abs(rng.normal(0, 30, (36, 36, 4))).astype(np.float32)
    out["lcn__img"] = img
    out["lcn__lcn_gpu"] = np.asarray(ref_pre.lcn_gpu(img.copy(), 5, (27, 27, 1)))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    np.savez_compressed(args.out, **out)
    print("wrote", args.out, os.path.getsize(args.out), "bytes, TensorFlow", tf.__version__)


if __name__ == "__main__":
    main()
