class DataLoader:  # import-only stub
    pass
