from . import agent_classes, agent_methods, predicates  # noqa: F401
