# In drop-in mode (this package's directory ahead of the reference's on sys.path) the modules that are not
# replaced here (the generators, the photometric / geometric operations ...) keep resolving to the reference's own
# package of the same name further down sys.path.
from pkgutil import extend_path
__path__ = extend_path(__path__, __name__)
