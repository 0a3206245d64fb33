"""Drop-in package with the reference's dotted name.  Modules this repo replaces live here; anything else the
reference keeps in its own package of the same name (e.g. ``data_load_preprocess.load_record``, network I/O) is
found further down ``sys.path`` -- the reference tree stays on the path BEHIND ``dropin/``."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
