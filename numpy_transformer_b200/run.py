"""Launcher: run one of the reference's scripts, unchanged, on the B200 backend.

    python -m numpy_transformer_b200.run [--device K] [--seed S] path/to/transformer_of_image.py [script args]

It does what a user of the reference would otherwise put in a sitecustomize: initialise the device, put the
drop-in mirror of the reference's modules (net.*, attention, classify, PatchEmbed, Position_add, gpt.*) first on
sys.path, then execute the script as __main__ (transformer_of_image.py:211, gpt/gpt_train_potry3000.py:374,
gpt_character/gpt_train_english_char.py:310).  The script's own directory tree stays importable for everything the
mirror does not provide.  `--seed` seeds np.random before the script runs (the scripts never seed; parity runs do).
"""
import os
import runpy
import sys


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    device, seed = 0, None
    while argv and argv[0].startswith("--"):
        opt = argv.pop(0)
        if opt == "--device":
            device = int(argv.pop(0))
        elif opt == "--seed":
            seed = int(argv.pop(0))
        else:
            raise SystemExit("unknown option %s" % opt)
    if not argv:
        raise SystemExit(__doc__)
    script = os.path.abspath(argv[0])
    import numpy as np
    import numpy_transformer_b200 as nt
    nt.init(device)
    nt.install()
    if seed is not None:
        np.random.seed(seed)
    sys.argv = [script] + argv[1:]
    try:
        runpy.run_path(script, run_name="__main__")
    finally:
        nt.synchronize()


if __name__ == "__main__":
    main()
