"""state_dict keys + shapes of the reference's networks, so that checkpoint compatibility of
fots.pytorch_b200.pipeline.nets can be checked without the reference present.

Run in the BUILD container:  python tests/golden/make_model_keys_golden.py
Imports /root/reference/tools/models.py unmodified (CPU); writes tests/golden/ref_model_keys.json.
"""
import importlib.util
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference_models():
    spec = importlib.util.spec_from_file_location("ref_models", "/root/reference/tools/models.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    ref = load_reference_models()
    out = {}
    for name, net in (("ModelResNetSep2_att_89", ref.ModelResNetSep2(attention=True, nclass=89)),
                      ("ModelResNetSep2_noatt_89", ref.ModelResNetSep2(attention=False, nclass=89)),
                      ("CRNN_89", ref.CRNN(nclass=89))):
        out[name] = {k: list(v.shape) for k, v in net.state_dict().items()}
        print(name, len(out[name]), "tensors", sum(v.numel() for v in net.state_dict().values()), "values")
    with open(os.path.join(HERE, "ref_model_keys.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
