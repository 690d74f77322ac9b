"""Config2D / Config3D.

Mirrors stardist/models/model2d.py:123-269 and model3d.py:129-311 (attribute names, defaults,
derived fields, JSON round trip) on top of a small restatement of csbdeep's BaseConfig
(csbdeep is an un-vendored dependency of the reference, setup.py:140).  Training-only fields are
kept so that a reference `config.json` loads unchanged; they are not used on this path.
"""
import argparse, json
import numpy as np
from ..utils import _normalize_grid, axes_check_and_normalize, _raise


class BaseConfig(argparse.Namespace):
    def __init__(self, axes='YX', n_channel_in=1, n_channel_out=1, allow_new_parameters=False, **kwargs):
        axes = axes_check_and_normalize(axes)
        ax = {a: axes.find(a) != -1 for a in 'STCZYX'}
        (ax['X'] and ax['Y']) or _raise(ValueError('lateral axes X and Y must be present.'))
        not (ax['Z'] and ax['T']) or _raise(ValueError('using Z and T axes together not supported.'))
        axes.startswith('S') or (not ax['S']) or _raise(ValueError('sample axis S must be first.'))
        axes = axes.replace('S', '')
        n_dim = 3 if (ax['Z'] or ax['T']) else 2
        # channels last
        if ax['C']:
            axes[-1] == 'C' or _raise(ValueError('channel axis must be last for backend.'))
        else:
            axes += 'C'
        self.n_dim = n_dim
        self.axes = axes
        self.n_channel_in = int(max(1, n_channel_in))
        self.n_channel_out = int(max(1, n_channel_out))
        self.train_checkpoint = 'weights_best.h5'
        self.train_checkpoint_last = 'weights_last.h5'
        self.train_checkpoint_epoch = 'weights_now.h5'
        self.update_parameters(allow_new_parameters, **kwargs)

    def is_valid(self, return_invalid=False):
        return (True, tuple()) if return_invalid else True

    def update_parameters(self, allow_new=False, **kwargs):
        if not allow_new:
            attr_new = [k for k in kwargs if not hasattr(self, k)]
            if len(attr_new) > 0:
                raise AttributeError("Not allowed to add new parameters (%s)" % ', '.join(attr_new))
        for k in kwargs:
            setattr(self, k, kwargs[k])

    def to_json(self):
        def _conv(v):
            if isinstance(v, tuple): return [_conv(x) for x in v]
            if isinstance(v, (np.integer,)): return int(v)
            if isinstance(v, (np.floating,)): return float(v)
            return v
        return json.dumps({k: _conv(v) for k, v in vars(self).items()})


class Config2D(BaseConfig):
    """Configuration for a :class:`StarDist2D` model (stardist/models/model2d.py:123-269)."""

    def __init__(self, axes='YX', n_rays=32, n_channel_in=1, grid=(1, 1), n_classes=None, backbone='unet', **kwargs):
        super().__init__(axes=axes, n_channel_in=n_channel_in, n_channel_out=1 + n_rays)
        self.n_rays = int(n_rays)
        self.grid = _normalize_grid(grid, 2)
        self.backbone = str(backbone).lower()
        self.n_classes = None if n_classes is None else int(n_classes)
        if self.backbone == 'unet':
            self.unet_n_depth = 3
            self.unet_kernel_size = 3, 3
            self.unet_n_filter_base = 32
            self.unet_n_conv_per_depth = 2
            self.unet_pool = 2, 2
            self.unet_activation = 'relu'
            self.unet_last_activation = 'relu'
            self.unet_batch_norm = False
            self.unet_dropout = 0.0
            self.unet_prefix = ''
            self.net_conv_after_unet = 128
        else:
            raise ValueError("backbone '%s' not supported." % self.backbone)
        self.net_input_shape = None, None, self.n_channel_in
        self.net_mask_shape = None, None, 1
        self.train_shape_completion = False
        self.train_completion_crop = 32
        self.train_patch_size = 256, 256
        self.train_background_reg = 1e-4
        self.train_foreground_only = 0.9
        self.train_sample_cache = True
        self.train_dist_loss = 'mae'
        self.train_loss_weights = (1, 0.2) if self.n_classes is None else (1, 0.2, 1)
        self.train_class_weights = (1, 1) if self.n_classes is None else (1,) * (self.n_classes + 1)
        self.train_epochs = 400
        self.train_steps_per_epoch = 100
        self.train_learning_rate = 0.0003
        self.train_batch_size = 4
        self.train_n_val_patches = None
        self.train_tensorboard = True
        self.train_reduce_lr = {'factor': 0.5, 'patience': 40, 'min_delta': 0}
        self.use_gpu = False
        for k in ('n_dim', 'n_channel_out'):
            try: del kwargs[k]
            except KeyError: pass
        self.update_parameters(False, **kwargs)
        self.grid = _normalize_grid(self.grid, 2)
        if not len(self.train_loss_weights) == (2 if self.n_classes is None else 3):
            raise ValueError(f"train_loss_weights {self.train_loss_weights} not compatible with n_classes ({self.n_classes})")
        if not len(self.train_class_weights) == (2 if self.n_classes is None else self.n_classes + 1):
            raise ValueError(f"train_class_weights {self.train_class_weights} not compatible with n_classes ({self.n_classes})")


class Config3D(BaseConfig):
    """Configuration for a :class:`StarDist3D` model (stardist/models/model3d.py:129-311)."""

    def __init__(self, axes='ZYX', rays=None, n_channel_in=1, grid=(1, 1, 1), n_classes=None, anisotropy=None, backbone='unet', **kwargs):
        from ..rays3d import Rays_GoldenSpiral, rays_from_json
        if rays is None:
            if 'rays_json' in kwargs:
                rays = rays_from_json(kwargs['rays_json'])
            elif 'n_rays' in kwargs:
                rays = Rays_GoldenSpiral(kwargs['n_rays'])
            else:
                rays = Rays_GoldenSpiral(96)
        elif np.isscalar(rays):
            rays = Rays_GoldenSpiral(rays)
        super().__init__(axes=axes, n_channel_in=n_channel_in, n_channel_out=1 + len(rays))
        self.n_rays = len(rays)
        self.grid = _normalize_grid(grid, 3)
        self.anisotropy = anisotropy if anisotropy is None else tuple(anisotropy)
        self.backbone = str(backbone).lower()
        self.rays_json = rays.to_json()
        self.n_classes = None if n_classes is None else int(n_classes)
        if 'anisotropy' in self.rays_json['kwargs']:
            if self.rays_json['kwargs']['anisotropy'] is None and self.anisotropy is not None:
                self.rays_json['kwargs']['anisotropy'] = self.anisotropy
                print("Changing 'anisotropy' of rays to %s" % str(anisotropy))
            elif self.rays_json['kwargs']['anisotropy'] is not None and self.anisotropy is not None and \
                    tuple(self.rays_json['kwargs']['anisotropy']) != tuple(self.anisotropy):
                import warnings
                warnings.warn("Mismatch of 'anisotropy' of rays and 'anisotropy'.")
        if self.backbone == 'unet':
            self.unet_n_depth = 2
            self.unet_kernel_size = 3, 3, 3
            self.unet_n_filter_base = 32
            self.unet_n_conv_per_depth = 2
            self.unet_pool = 2, 2, 2
            self.unet_activation = 'relu'
            self.unet_last_activation = 'relu'
            self.unet_batch_norm = False
            self.unet_dropout = 0.0
            self.unet_prefix = ''
            self.net_conv_after_unet = 128
        elif self.backbone == 'resnet':      # model3d.py:256-264
            self.resnet_n_blocks = 4
            self.resnet_kernel_size = 3, 3, 3
            self.resnet_kernel_init = 'he_normal'
            self.resnet_n_filter_base = 32
            self.resnet_n_conv_per_block = 3
            self.resnet_activation = 'relu'
            self.resnet_batch_norm = False
            self.net_conv_after_resnet = 128
        else:
            raise ValueError("backbone '%s' not supported." % self.backbone)
        self.net_input_shape = None, None, None, self.n_channel_in
        self.net_mask_shape = None, None, None, 1
        self.train_patch_size = 128, 128, 128
        self.train_background_reg = 1e-4
        self.train_foreground_only = 0.9
        self.train_sample_cache = True
        self.train_dist_loss = 'mae'
        self.train_loss_weights = (1, 0.2) if self.n_classes is None else (1, 0.2, 1)
        self.train_class_weights = (1, 1) if self.n_classes is None else (1,) * (self.n_classes + 1)
        self.train_epochs = 400
        self.train_steps_per_epoch = 100
        self.train_learning_rate = 0.0003
        self.train_batch_size = 1
        self.train_n_val_patches = None
        self.train_tensorboard = True
        self.train_reduce_lr = {'factor': 0.5, 'patience': 40, 'min_delta': 0}
        self.use_gpu = False
        for k in ('n_dim', 'n_channel_out', 'n_rays', 'rays_json'):
            try: del kwargs[k]
            except KeyError: pass
        self.update_parameters(False, **kwargs)
        self.grid = _normalize_grid(self.grid, 3)
