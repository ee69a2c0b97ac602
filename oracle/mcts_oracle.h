/* placeholder - filled in with the search oracle */
