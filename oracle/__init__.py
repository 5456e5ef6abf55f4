"""oracle/ — CPU restatement of the ZouJiu1/numpy_transformer training step.

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import this package; the product
path (numpy_transformer_b200/) never does and fails loudly without its CUDA library.

Parity status: PINNED.  oracle/gen_golden.py imports the unmodified reference from
/root/reference and writes tests/golden/*.npz; tests/test_oracle_golden.py checks every
oracle function and both model-level steps against those vectors (<= 1e-12 rel, fp64).
"""
