"""Empty stand-in so the reference's `import matplotlib.pyplot` (Position_add.py:2,
gpt/FixedEmbed.py:9) resolves when oracle/gen_golden.py imports it. Test infrastructure only."""
