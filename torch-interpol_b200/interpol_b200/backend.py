"""Backend switches.  `jitfields` mirrors the reference's flag
(`interpol/backend.py:1`) for source compatibility; it has no effect here: the
CUDA extension is the only backend."""
jitfields = False
