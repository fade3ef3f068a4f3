"""geometric multigrid (device) -- see hierarchy.py"""
