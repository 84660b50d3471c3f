from . import functional, init
from .modules import *
from .parameter import Parameter
