"""Reference arm: runs the UNMODIFIED upstream repository (copied to ``baseline/_ref`` by
``baseline/install_reference.py``; git-ignored, travels to the GPU box) with stand-ins for its missing
third-party imports (``baseline/stubs.py``).  Bench / test infrastructure only: nothing under
``relation-detr_b200/`` imports this package."""
