"""Model registry - mirror of torchsr/models.py:10-82 (same names, return values and error behaviour)."""
from argparse import Namespace
from typing import Tuple

from .esrgan.generator import Generator as ESRGANGenerator
from .esrgan.trainer import ESRGANTrainer
from .srgan.generator import Generator as SRGANGenerator
from .srgan.trainer import SRGANTrainer

MODELS = {'esrgan': ESRGANTrainer, 'srgan': SRGANTrainer}
CROP_SIZE = {'esrgan': 128, 'srgan': 96}
GENERATORS = {'esrgan': ESRGANGenerator, 'srgan': SRGANGenerator}


def select_trainer_model(args: Namespace) -> Tuple[object, int]:
    """Returns (trainer class, HR crop size) for args.model (case-insensitive); RuntimeError for unknown names."""
    model = args.model.lower()
    try:
        return MODELS[model], CROP_SIZE[model]
    except KeyError:
        raise RuntimeError(f'Unknown model: {model}. Available models: {MODELS.keys()}')


def select_test_model(args: Namespace) -> object:
    """Returns the generator class for args.model; RuntimeError for unknown names."""
    model = args.model.lower()
    try:
        return GENERATORS[model]
    except KeyError:
        raise RuntimeError(f'Unknown model: {model}. Available models: {GENERATORS.keys()}')
