"""Naming surface shared by the filter builders: :class:`StrategyDict`.

The reference addresses alternative designs of one filter as ``lowpass.pole``,
``gammatone["slaney"]``, ``erb.gm90`` ... through its ``StrategyDict``
(reference ``audiolazy/lazy_core.py:431-659``). Only that addressing scheme is
re-created here (the reference's metaclass/operator machinery is out of scope).
"""
from __future__ import annotations


class StrategyDict(object):
  """Callable collection of named implementations ("strategies") of one function.

  * ``sd.strategy("name", "alias", ...)`` is a decorator registering a function under
    every given name; it returns the :class:`StrategyDict` itself, so the usual idiom
    ``@sd.strategy("x")\\ndef sd(...): ...`` keeps ``sd`` bound to the collection.
  * ``sd["name"]`` and ``sd.name`` give the strategy; ``sd(...)`` calls ``sd.default``
    (the first one registered unless reassigned, like ``lowpass.default = lowpass.pole``
    in reference ``lazy_filters.py:1494-1495``).
  * iterating yields each distinct strategy once, in registration order.
  """

  def __init__(self, name="strategy_dict_unnamed_instance"):
    object.__setattr__(self, "_names", {})      # name -> function
    object.__setattr__(self, "_unique", [])     # functions, registration order
    object.__setattr__(self, "__name__", name)
    object.__setattr__(self, "default", None)

  # -- registration ------------------------------------------------------------------
  def strategy(self, *names):
    if not names:
      raise TypeError("a strategy needs at least one name")

    def register(func):
      func.__name__ = str(names[0])
      for name in names:
        self._names[name] = func
      self._unique.append(func)
      if self.default is None:
        object.__setattr__(self, "default", func)
      return self

    return register

  # -- access ------------------------------------------------------------------------
  def __getitem__(self, name):
    if isinstance(name, tuple):   # the reference's MultiKeyDict accepts the alias tuple
      name = name[0]
    return self._names[name]

  def __getattr__(self, name):
    names = object.__getattribute__(self, "_names")
    if name in names:
      return names[name]
    raise AttributeError("%r has no strategy %r" % (object.__getattribute__(self, "__name__"), name))

  def __setattr__(self, name, value):
    if name == "default" or name.startswith("_"):
      object.__setattr__(self, name, value)
    else:
      self._names[name] = value
      if value not in self._unique:
        self._unique.append(value)

  def __contains__(self, name):
    return name in self._names

  def __call__(self, *args, **kwargs):
    if self.default is None:
      raise NotImplementedError("StrategyDict %r has no strategy" % self.__name__)
    return self.default(*args, **kwargs)

  def __iter__(self):
    return iter(list(self._unique))

  def __len__(self):
    return len(self._unique)

  # -- the dictionary view of the reference's MultiKeyDict: one entry per strategy, keyed by the tuple of its names
  def value2keys(self, func):
    """All names of a strategy, in registration order (``()`` when it is not one)."""
    return tuple(name for name, value in self._names.items() if value is func)

  def key2keys(self, name):
    """All names of the strategy that ``name`` designates."""
    return self.value2keys(self._names[name]) if name in self._names else ()

  def keys(self):
    return [self.value2keys(func) for func in self._unique]

  def values(self):
    return list(self._unique)

  def items(self):
    return [(self.value2keys(func), func) for func in self._unique]

  def __repr__(self):
    return "<StrategyDict %s: %s>" % (self.__name__, ", ".join(f.__name__ for f in self._unique))
