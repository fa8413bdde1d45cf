"""The path-relevant configuration keys of the reference, under the reference's own names.

The reference keeps one flat YAML per experiment (`self.config.get(...)`); the keys that reach the
per-step part-disentanglement path are
    batch_size, spatial_size, n_parts, local_app_size, use_tps, tps_parameters{scal, tps_scal,
    rot_scal, off_scal, scal_var, augm_scal}
(cub/code/SB_model48i/train_cub_subset_tps.yaml:19-20,132,139,141,187-194).  `PathConfig` carries
exactly those; the presets are the shipped / logged values of the three datasets.
"""
from dataclasses import dataclass, field, asdict

# cub/code/SB_model48i/train_cub_subset_tps.yaml:188-194
CUB_TPS = dict(scal=0.8, tps_scal=0.15, rot_scal=0.2, off_scal=0.2, scal_var=0.1, augm_scal=1.0)
# the use_tps run of PennAction, pennaction/log.txt:2013-2019
PENN_TPS = dict(scal=0.95, tps_scal=0.08, rot_scal=0.05, off_scal=0.2, scal_var=0.05, augm_scal=1.0)


@dataclass
class PathConfig:
    batch_size: int = 8
    spatial_size: int = 128
    n_parts: int = 25
    local_app_size: int = 64
    use_tps: bool = True
    n_views: int = 3                      # CUB: view0, view1, view0_target (model.py:298-310); PennAction / DeepFashion: 2
    tps_parameters: dict = field(default_factory=lambda: dict(CUB_TPS))

    @classmethod
    def from_dict(cls, cfg):
        """Pick the path's keys out of a reference-style config dict (unknown keys are ignored)."""
        names = {f for f in cls.__dataclass_fields__}
        return cls(**{k: v for k, v in dict(cfg).items() if k in names})

    def to_dict(self):
        return asdict(self)

    def make_step(self, **kw):
        """PartStep for this configuration."""
        from .step import PartStep
        return PartStep(self.batch_size, self.spatial_size, self.n_parts, self.local_app_size, n_views=self.n_views,
                        use_tps=self.use_tps, **kw)


# shipped training shapes and the benchmark shapes of BASELINE.json
CUB_SHIPPED = PathConfig(8, 128, 25, 64, True, 3, dict(CUB_TPS))                 # train_cub_subset_tps.yaml:19-20,132,139
CUB_BENCH = PathConfig(256, 128, 16, 64, True, 3, dict(CUB_TPS))                 # BASELINE.json configs[1]
DEEPFASHION_BENCH = PathConfig(128, 256, 16, 64, True, 2, dict(PENN_TPS))        # configs[2]; 256^2: deepfashion/code/iccv19/train.yaml:12
#   (the reference's DeepFashion model has no TPS, deepfashion/code/SB_model48c/model.py:266-280; SURVEY.md 8d config 3 benches V=2)
PENNACTION_BENCH = PathConfig(512, 128, 16, 64, True, 2, dict(PENN_TPS))         # configs[3]; pennaction/log.txt:2013-2019
