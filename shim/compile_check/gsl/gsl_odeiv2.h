/* syntax-check stub */
