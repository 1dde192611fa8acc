__all__ = ["Graph"]
class Graph:  # import-only stub
    pass
