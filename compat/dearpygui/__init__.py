"""stand-in for dearpygui (GUI only)"""
