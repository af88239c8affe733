"""Host-side mirror of the reference plugin registry.

The reference package cannot be imported under python 3.12 (zeus/common/utils.py:17 `import imp`),
so this module restates the two pieces of it the hot-path plugins depend on:

* ``Registers``    -- zeus/common/util/register.py:39-83: ``@Registers.model`` /
  ``.algorithm`` / ``.agent`` / ``.env`` / ``.comm`` store a class under ``cls.__name__``;
  re-registering replaces the entry with a warning; lookup of a missing key logs and re-raises.
* ``import_config`` -- zeus/common/util/common.py:32-44: overwrite same-named module globals.

When the real ``zeus`` package IS importable (a maintainer's environment), ``install()`` registers
the B200 classes into the reference's own ``Registers`` so ``xt_main`` picks them up
(INTEGRATION.md)."""
import logging


class RegisterStub(object):
    """zeus/common/util/register.py:39-70."""

    def __init__(self, name):
        self._dict = dict()
        self._name = name

    def __getitem__(self, key):
        try:
            return self._dict[key]
        except Exception as exc:
            logging.error("module %s not found: %s", key, exc)
            raise exc

    def __contains__(self, key):
        return key in self._dict

    def __call__(self, param):
        if not callable(param):
            raise Exception("To Registry must be callable, Got: {}.".format(param))
        register_name = param.__name__
        if register_name in self._dict:
            logging.warning("Key:%s is registered, will replace with %s.", register_name, self._name)
        self._dict[register_name] = param
        return param

    def keys(self):
        return self._dict.keys()


class Registers(object):  # pylint: disable=too-few-public-methods
    """zeus/common/util/register.py:73-83."""

    def __init__(self):
        raise RuntimeError("Registries prohibit instancing !")

    agent = RegisterStub("agent")
    model = RegisterStub("model")
    algorithm = RegisterStub("algorithm")
    env = RegisterStub("env")
    comm = RegisterStub("comm")


def import_config(global_para, config):
    """zeus/common/util/common.py:32-44."""
    if not config:
        return
    for key in config.keys():
        if key in global_para:
            global_para[key] = config[key]


def model_builder(model_info):
    """xt/model/__init__.py:15-19."""
    return Registers.model[model_info["model_name"]](model_info)


def alg_builder(alg_name, model_info, alg_config, **kwargs):
    """xt/algorithm/__init__.py:19-28."""
    return Registers.algorithm[alg_name](model_info=model_info, alg_config=alg_config, **kwargs)


def install(target_registers=None):
    """Re-register every B200 plugin into the reference's ``Registers`` (same keys, so the
    reference's ``@Registers`` replacement-with-warning semantics make them the active ones)."""
    if target_registers is None:
        from zeus.common.util.register import Registers as target_registers  # noqa: only in a reference env
    for kind in ("model", "algorithm", "agent", "comm"):
        src = getattr(Registers, kind)
        dst = getattr(target_registers, kind, None)
        if dst is None:
            continue
        for key in list(src.keys()):
            dst(src[key])
    return target_registers
