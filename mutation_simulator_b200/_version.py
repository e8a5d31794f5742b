__version__ = "3.0.2+b200.1"
REFERENCE_VERSION = "3.0.2"
