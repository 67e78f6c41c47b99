"""`miscc` of the drop-in: config and losses resolve to the B200 implementation; every other submodule (miscc.utils: the
attention visualisation, mkdir_p, ...) falls through to a reference `miscc/` directory further down sys.path."""
import pkgutil

__path__ = pkgutil.extend_path(__path__, __name__)
