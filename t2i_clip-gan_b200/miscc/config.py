"""The slice of the reference's global ``cfg`` (DMGAN+CLIP/code/miscc/config.py:9-78) that the hot path reads:
``cfg.CUDA`` (losses.py:65,259) and ``cfg.TRAIN.SMOOTH.GAMMA1/2/3, LAMBDA`` (losses.py:79; yml files)."""


class _Node(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


cfg = _Node(CUDA=True,
            TRAIN=_Node(SMOOTH=_Node(GAMMA1=4.0, GAMMA2=5.0, GAMMA3=10.0, LAMBDA=10.0)))
