"""Model Context Protocol server: five consolidated tools + the legacy tool names, over stdio or streamable HTTP."""
