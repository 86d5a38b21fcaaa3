"""`simulation.log` helpers, API and line formats of reference
`src/phasefieldx/Logger/library_versions.py:23-130`; version lines tolerate a missing dolfinx."""
import logging
import os
import platform
import sys
import time

import numpy as np


def set_logger(result_folder_name):
    logger = logging.getLogger("simulation_logger")
    for handler in list(logger.handlers):
        logger.removeHandler(handler)
        handler.close()
    logger.setLevel(logging.INFO)
    fh = logging.FileHandler(os.path.join(result_folder_name, "simulation.log"))
    fh.setFormatter(logging.Formatter("%(message)s"))
    logger.addHandler(fh)
    return logger


def _version(mod):
    try:
        return __import__(mod).__version__
    except Exception:
        return "not installed"


def log_library_versions(logger):
    import phasefieldx_b200
    logger.info(f"Python version: {sys.version_info.major}.{sys.version_info.minor}")
    logger.info("=========== Library Versions ===========")
    logger.info(f"PhaseFieldX : {phasefieldx_b200.__version__}")
    logger.info(f"DolfinX : {_version('dolfinx')}")
    logger.info(f"ufl : {_version('ufl')}")
    logger.info(f"basix : {_version('basix')}")
    logger.info(f"numpy : {np.__version__}")
    logger.info(f"logging : {logging.__version__}")
    logger.info("=======================================")


def log_system_info(logger):
    logger.info("=========== Platform ==================")
    logger.info(f"Operating System Information: {platform.platform()}")
    logger.info(f"Architecture : {platform.architecture()}")
    logger.info(f"User name : {platform.uname()}")
    logger.info(f"processor : {platform.processor()}")
    logger.info(f"Machine type : {platform.machine()}")
    logger.info(f"Python version : {platform.python_version()}")
    logger.info("=======================================")


def log_end_analysis(logger, totaltime=0.0):
    logger.info("\n\n\n ====================================================")
    logger.info("\n\n End of computations")
    logger.info(" Analysis finished correctly.")
    logger.info(f" total simulation time: {totaltime}")
    logger.info(f"Analysis finished on {time.strftime('%a %b %d %H:%M:%S %Y', time.localtime())}")


def log_model_information(msh, logger):
    logger.info("========== Mesh Information ===========")
    logger.info(f"Dimension: {msh.geometry.dim}")
