"""Reference arm of bench.py: the UNMODIFIED reference installed under baseline/_ref/ (git-ignored,
travels to the GPU box with the snapshot).  See install_ref.sh / ref_loader.py."""
