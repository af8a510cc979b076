"""rustracer_b200 — B200-native wavefront renderer behind rustracer's Scene / SamplerIntegrator API.

Only what the hot path needs lives here: the host front end (`host`), the device binding (`device`),
the integrator mirror (`integrator`) and the synthetic scene generators for the benchmark configs (`scenes`).
"""
from .host import Scene, SceneError  # noqa: F401
