import pkgutil

__path__ = pkgutil.extend_path(__path__, __name__)      # utils.experiments, utils.logging_config: the reference's own files
