# flake8: noqa
"""Mirror of diffsci.models.karras (reference karras/__init__.py) for the EDM hot path."""
from .karrasmodule import KarrasModule, KarrasModuleConfig
from .karrasmodule_new import EnsembleKarrasModule, EnsembleKarrasModuleConfig
from .schedulers import Scheduler, EDMScheduler, VPScheduler, VEScheduler
from .noisesamplers import NoiseSampler, EDMNoiseSampler, VPNoiseSampler, VENoiseSampler, UniformNoiseSampler
from .schedulingfunctions import (SchedulingFunctions, EDMSchedulingFunctions, VPSchedulingFunctions,
                                  VESchedulingFunctions)
from .preconditioners import (KarrasPreconditioner, EDMPreconditioner, VPPreconditioner, VEPreconditioner,
                              NullPreconditioner, SR3Preconditioner)
from .integrators import (Integrator, EulerIntegrator, HeunIntegrator, EulerMaruyamaIntegrator, KarrasIntegrator,
                          name_to_integrator)
from .ema import ModelEMA
from .engine import SamplerEngine
from .trainer import EDMTrainer
