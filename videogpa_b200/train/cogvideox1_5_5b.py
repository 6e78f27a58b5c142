"""`train/CogVideoX1.5-5B/03_train.py` of the reference: the CogVideoX-5B trainer (train/cogvideox_5b.py) on the temporal-patch
transformer.

Reference differences to train/CogVideoX-5B/03_train.py, all kept: DEFAULT_CONFIG (`model_path THUDM/CogVideoX1.5-5B`,
`max_steps 1500`, experiment name; :37-75), no VAE in the module (latents are precomputed, :87-115), latents trimmed to even
F / H / W before noising (:134-144, done by train_step.DPOSharedStep for patch_size_t models). `dataset15` imported there
(:28) is not in the repository; the on-disk format is the one `train/dataset.py` reads (videogpa_b200.dataset).
`use_dynamic_positional_embedding` (:95,112) concerns the rotary table, which the training call does not pass.
"""
from __future__ import annotations

from . import cogvideox_5b as base

DEFAULT_CONFIG = dict(base.DEFAULT_CONFIG)
DEFAULT_CONFIG.update({"model_path": "THUDM/CogVideoX1.5-5B", "max_steps": 1500, "experiment_name": "cogvideo1.5_dpo_t2v"})
for _k in ("enable_slicing", "enable_tiling"):          # the 1.5 module has no VAE
    DEFAULT_CONFIG.pop(_k, None)


def load_config(args) -> dict:
    """DEFAULT_CONFIG <- YAML `training:` section <- --devices (reference :245-262)."""
    config = dict(DEFAULT_CONFIG)
    config["base_path"] = args.base_path
    if args.config:
        import yaml
        with open(args.config, "r") as f:
            config.update(yaml.safe_load(f).get("training", {}))
    if args.devices:
        config["devices"] = [int(d) for d in args.devices.split(",")]
    config.setdefault("synthetic_variant", "1.5")
    return config


def main(argv=None):
    args = base.build_parser().parse_args(argv)
    return base.main_train(load_config(args), synthetic_layers=args.synthetic)


if __name__ == "__main__":
    main()
