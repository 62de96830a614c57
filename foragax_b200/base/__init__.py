from . import agent_classes, agent_methods, predicates, space_classes, space_methods  # noqa: F401
