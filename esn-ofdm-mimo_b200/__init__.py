# The directory name carries a hyphen (it mirrors the reference repository's
# name) and is therefore not importable as a package; put this directory and
# `libs/` on sys.path instead.
