"""triplaneturbo_b200 — B200-native (sm_100a) triplane volume-rendering hot path behind the threestudio /
triplaneturbo_executable renderer + geometry plugin API.  See DESIGN.md; C ABI in include/triplane_b200.h."""
from .compat import register, find, C, BaseModule  # noqa: F401
from . import geometry, renderer, nerfacc_compat, sampler  # noqa: F401  (registers the plugins)
from .sampler import grid_sample, grid_sample_2d, sample_from_planes, project_onto_planes  # noqa: F401
from .geometry import StableDiffusionTriplaneDualAttention, VanillaMLP  # noqa: F401
from .renderer import (GenerativeSpaceSDFVolumeRenderer, PatchRenderer, ImportanceEstimator, LearnedVariance,  # noqa: F401
                       NoMaterial, SolidColorBackground)

__all__ = ["register", "find", "StableDiffusionTriplaneDualAttention", "GenerativeSpaceSDFVolumeRenderer",
           "PatchRenderer", "ImportanceEstimator", "NoMaterial", "SolidColorBackground", "grid_sample", "grid_sample_2d",
           "sample_from_planes", "project_onto_planes"]
