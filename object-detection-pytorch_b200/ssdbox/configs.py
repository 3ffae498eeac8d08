"""Named box-path configurations (the five rows of BASELINE.json ``configs``).

Each entry carries the ``cfg.MODEL`` prior fields the reference reads
(lib/utils/config.py:116-124; cfgs/vgg/ssd_vgg16_voc_image512.yml:13-23;
lib/models/rfb_net.py:344-347; lib/datasets/config.py:99-115) plus the feature-map sizes the
reference obtains from forward hooks (lib/models/__init__.py:37-54) and the batch / class
counts the metric is quoted on.  RefineDet320 is not in the reference snapshot; its prior
layout follows arXiv 1711.06897 (4 maps, 3 anchors per cell).
"""


class AttrDict(dict):
    """Minimal stand-in for lib/utils/config.py:16-31 (attribute + item access)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value


def _model(image, steps, mins, maxs, ars, num_classes, clip=True, flip=True):
    return AttrDict(IMAGE_SIZE=(image, image), STEPS=list(steps), MIN_SIZES=list(mins),
                    MAX_SIZES=list(maxs), ASPECT_RATIOS=[list(a) for a in ars],
                    VARIANCE=[0.1, 0.2], CLIP=clip, FLIP=flip, NUM_CLASSES=num_classes)


CONFIGS = {
    # cfg1: SSD300 VGG16 VOC, P=8732, C=21, B=32
    "ssd300_voc": dict(
        model=_model(300, [8, 16, 32, 64, 100, 300], [30, 60, 111, 162, 213, 264],
                     [60, 111, 162, 213, 264, 315], [[2], [2, 3], [2, 3], [2, 3], [2], [2]], 21),
        layer_dims=[[38, 38], [19, 19], [10, 10], [5, 5], [3, 3], [1, 1]],
        batch=32, gt_max=16, num_priors=8732),
    # cfg2: SSD512 VGG16 COCO, P=24564, C=81, B=64 (north-star workload)
    "ssd512_coco": dict(
        model=_model(512, [8, 16, 32, 64, 128, 256, 512],
                     [20.48, 51.2, 133.12, 215.04, 296.96, 378.88, 460.8],
                     [51.2, 133.12, 215.04, 296.96, 378.88, 460.8, 542.72],
                     [[2], [2, 3], [2, 3], [2, 3], [2, 3], [2], [2]], 81),
        layer_dims=[[64, 64], [32, 32], [16, 16], [8, 8], [4, 4], [2, 2], [1, 1]],
        batch=64, gt_max=32, num_priors=24564),
    # cfg3: RFBNet300 VGG16 VOC inference, P=11620, C=21, B=256
    "rfb300_voc": dict(
        model=_model(300, [8, 16, 32, 64, 100, 300], [30, 60, 111, 162, 213, 264],
                     [60, 111, 162, 213, 264, 315], [[2, 3], [2, 3], [2, 3], [2, 3], [2], [2]], 21),
        layer_dims=[[38, 38], [19, 19], [10, 10], [5, 5], [3, 3], [1, 1]],
        batch=256, gt_max=16, num_priors=11620),
    # cfg4: FSSD300 VGG COCO training targets, P=8732, C=81, B=32
    "fssd300_coco": dict(
        model=_model(300, [8, 16, 32, 64, 100, 300], [21, 45, 99, 153, 207, 261],
                     [45, 99, 153, 207, 261, 315], [[2], [2, 3], [2, 3], [2, 3], [2], [2]], 81),
        layer_dims=[[38, 38], [19, 19], [10, 10], [5, 5], [3, 3], [1, 1]],
        batch=32, gt_max=32, num_priors=8732),
    # cfg5: RefineDet320 VOC, P=6375, ARM C=2 / ODM C=21, B=32 (parity unpinned)
    "refinedet320_voc": dict(
        model=_model(320, [8, 16, 32, 64], [32, 64, 128, 256], [], [[2], [2], [2], [2]], 21),
        layer_dims=[[40, 40], [20, 20], [10, 10], [5, 5]],
        batch=32, gt_max=16, num_priors=6375),
}


def get(name):
    c = CONFIGS[name]
    return AttrDict(MODEL=c["model"]), c
