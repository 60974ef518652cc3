/* nothing of gsl_sf is used */
