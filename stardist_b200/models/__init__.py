from .config import Config2D
from .model2d import StarDist2D
from .base import StarDistPadAndCropResizer, PercentileNormalizer, NoNormalizer
