"""TEST INFRASTRUCTURE ONLY -- dumps the state_dict layout of the REAL reference ``cv_clf`` (VanillaClassifier with the ViT
encoder, cflearn/modules/cv/classifier/vanilla.py:16-66) built through the reference's own ``build_module``:
tests/golden/cv_clf_vit_tiny_keys.json.  Run in the build container (needs /root/reference)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from load_reference import load_modules  # noqa: E402


def main() -> None:
    mods = load_modules()
    m = mods.build_module("cv_clf", config=dict(in_channels=3, num_classes=10, img_size=32, latent_dim=128, encoder="vit",
                                                encoder_config=dict(patch_size=16, num_layers=2)))
    keys = [[k, list(v.shape)] for k, v in m.state_dict().items()]
    out = os.path.join(ROOT, "tests", "golden", "cv_clf_vit_tiny_keys.json")
    with open(out, "w") as f:
        json.dump({"generator": "oracle/make_golden_cv_clf.py (reference @ ca5ced1 through oracle/load_reference.py)",
                   "num_params": sum(p.numel() for p in m.parameters()), "keys": keys}, f, indent=1)
    print(f"wrote {out}: {len(keys)} keys")


if __name__ == "__main__":
    main()
