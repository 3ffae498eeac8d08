"""PriorBoxSSD -- drop-in for lib/layers/functions/prior_box.py (PriorBoxBase :20-111,
PriorBoxSSD :114-143).  Same constructor (`cfg` with cfg.MODEL.*), `.num_priors`, and
`.forward(layer_dims, tb_writer=None, image=None)`; the anchors come from the
`priorbox_kernel` of libssdbox.so (fp64 arithmetic, one rounding to fp32, optional clamp)."""
import torch

from . import _abi


def _field(model, key):
    if isinstance(model, dict):
        return model[key]
    return getattr(model, key)


def _has(model, key):
    if isinstance(model, dict):
        return key in model
    return hasattr(model, key)


class PriorBoxBase(object):
    def __init__(self, cfg):
        super(PriorBoxBase, self).__init__()
        model = cfg["MODEL"] if isinstance(cfg, dict) and "MODEL" in cfg else getattr(cfg, "MODEL", cfg)
        self._model = model
        self.image_size = _field(model, "IMAGE_SIZE")
        self._steps = _field(model, "STEPS")
        self._cfg_list = []
        self._prior_cfg = {}
        self._clip = _field(model, "CLIP")
        self._variance = _field(model, "VARIANCE")
        for v in self._variance:                       # prior_box.py:33-35
            if v <= 0:
                raise ValueError('Variances must be greater than 0')

    def _setup(self, cfg):
        num_feat = len(self._steps)
        for item in self._cfg_list:                    # prior_box.py:39-44
            if not _has(self._model, item):
                raise Exception("wrong anchor config!")
            val = _field(self._model, item)
            if len(val) != num_feat and len(val) != 0:
                raise Exception("config {} length does not match step length!".format(item))
            self._prior_cfg[item] = val


class PriorBoxSSD(PriorBoxBase):
    def __init__(self, cfg):
        super(PriorBoxSSD, self).__init__(cfg)
        self._cfg_list = ['MIN_SIZES', 'MAX_SIZES', 'ASPECT_RATIOS']
        self._flip = _field(self._model, "FLIP")
        self._setup(cfg)

    # ---- helpers -------------------------------------------------------------------------
    def _min_sizes(self, k):
        ms = self._prior_cfg['MIN_SIZES'][k]
        return list(ms) if isinstance(ms, (list, tuple)) else [ms]

    @property
    def num_priors(self):
        """priors per feature-map cell for every layer (prior_box.py:46-50), e.g. [4,6,6,6,4,4]."""
        has_max = len(self._prior_cfg['MAX_SIZES']) != 0
        out = []
        for k in range(len(self._steps)):
            per_min = 1 + (1 if has_max else 0) + len(self._prior_cfg['ASPECT_RATIOS'][k]) * (2 if self._flip else 1)
            out.append(len(self._min_sizes(k)) * per_min)
        return out

    def _abi_cfg(self, layer_dims):
        n = len(layer_dims)
        if n > _abi.MAX_LAYERS:
            raise Exception("too many feature maps: %d > %d" % (n, _abi.MAX_LAYERS))
        c = _abi.PriorCfg()
        c.num_layers = n
        c.clip = 1 if self._clip else 0
        c.flip = 1 if self._flip else 0
        has_max = len(self._prior_cfg['MAX_SIZES']) != 0
        c.has_max = 1 if has_max else 0
        c.image_h = float(self.image_size[0])
        c.image_w = float(self.image_size[1])
        for k in range(n):
            c.feat_h[k] = int(layer_dims[k][0])
            c.feat_w[k] = int(layer_dims[k][1])
            c.step[k] = float(self._steps[k])
            ms = self._min_sizes(k)
            ars = list(self._prior_cfg['ASPECT_RATIOS'][k])
            if len(ms) > _abi.MAX_MIN_SIZES or len(ars) > _abi.MAX_RATIOS:
                raise Exception("wrong anchor config!")
            c.num_min[k] = len(ms)
            for i, v in enumerate(ms):
                c.min_size[k][i] = float(v)
            if has_max:
                mx = self._prior_cfg['MAX_SIZES'][k]
                assert type(mx) is not list           # prior_box.py:134 one max size per layer
                c.max_size[k] = float(mx)
            c.num_ratio[k] = len(ars)
            for i, v in enumerate(ars):
                c.ratio[k][i] = float(v)
        return c

    def forward(self, layer_dims, tb_writer=None, image=None, device=None, keep_on_device=False):
        """Returns FloatTensor[P,4] (cx,cy,w,h).  Like the reference the result is a CPU tensor
        (train.py:63 moves it with .cuda()); pass keep_on_device=True to skip the round trip.
        tb_writer / image (visualisation, prior_box.py:55-90) are accepted and ignored."""
        c = self._abi_cfg(layer_dims)
        lib = _abi.lib()
        count = int(lib.ssdbox_priorbox_count(c))
        if count < 0:
            raise Exception("wrong anchor config! (%s)" % _abi.last_error())
        if not torch.cuda.is_available():
            raise RuntimeError("ssdbox: PriorBoxSSD.forward needs a CUDA device (no CPU path)")
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        out = torch.empty(count, 4, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _abi.check(lib.ssdbox_priorbox(c, _abi.ptr(out), count, _abi.stream_ptr(dev)))
        return out if keep_on_device else out.cpu()

    __call__ = forward
