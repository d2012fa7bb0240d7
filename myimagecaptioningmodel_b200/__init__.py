"""myimagecaptioningmodel_b200 -- the caption-decoder hot path of Chgtaxihe/MyImageCaptioningModel as hand-written
sm_100a CUDA behind a C-ABI (include/capdec.h), with the reference's Decoder / ImageCaptionModel interface on top."""
from ._capi import CapdecError, PARAM_NAMES, library_path, load_library
from .decoder import (AdaptorStore, CaptionDecoderEngine, Decoder, DecoderConfig, ImageCaptionModel, ParameterStore, default_engine,
                      caption_lengths, filter_caption, param_shapes)
from . import checkpoint, evaluation, synth
from .optim import Adam, LRSchedule, get_lr, get_optimizer
from .parallel import DataParallelTrainer, shard_range, sharded_decode

__all__ = ["checkpoint", "evaluation", "AdaptorStore", "Adam", "LRSchedule", "get_lr", "get_optimizer", "CapdecError", "PARAM_NAMES", "library_path", "load_library", "CaptionDecoderEngine", "Decoder",
           "DecoderConfig", "ImageCaptionModel", "ParameterStore", "default_engine", "caption_lengths", "filter_caption", "param_shapes",
           "DataParallelTrainer", "shard_range", "sharded_decode"]
