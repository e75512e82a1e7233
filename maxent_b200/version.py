"""Version report with the reference's two entry points (python/version.py.in:24-33)."""
version = "0.2"
reference_version = "1.2.0"          # the TRIQS/maxent release whose interface and numbers this package follows


def show_version():
    print("\nYou are using maxent_b200 version %s (interface of TRIQS/maxent %s)\n" % (version, reference_version))


def show_git_hash():
    print("\nmaxent_b200 is built in-tree; see `git log -1` of the checkout (no TRIQS dependency)\n")
