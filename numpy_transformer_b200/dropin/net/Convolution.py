"""net.Convolution — import-time placeholder.  The conv patch embed is reached only with
patch_convolu=1 (transformer_of_image.py:41 sets 0) and is out of the hot-path scope."""


class convolution_layer(object):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("convolution_layer is outside the B200 hot-path scope (SURVEY.md §8f rank 4)")
