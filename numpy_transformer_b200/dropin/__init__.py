"""Drop-in mirror of the reference's importable modules.

Put this directory (and dropin/gpt for the bare `attdecoderblock` spelling) on sys.path and the
reference's trainers import these classes instead of the NumPy ones:

    import numpy_transformer_b200; numpy_transformer_b200.install()
    from net.fullconnect import fclayer          # -> B200 backend
"""
