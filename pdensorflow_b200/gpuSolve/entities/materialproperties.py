"""MaterialProperties - mirror of gpuSolve/entities/materialproperties.py:8-199 (host-side container
of element / nodal properties and user-defined property functions; no device work)."""


class MaterialProperties:
    def __init__(self):
        self._element_properties = None
        self._nodal_properties = None
        self._ud_functions = None

    # ---- user-defined functions ----
    def add_ud_function(self, fname, fdef):
        if self._ud_functions is None:
            self._ud_functions = {}
        self._ud_functions[fname] = fdef

    def remove_ud_function(self, fname):
        if self._ud_functions is not None:
            self._ud_functions.pop(fname, None)
            if not self._ud_functions:
                self._ud_functions = None

    def execute_ud_func(self, fname, *kwargs):
        if self._ud_functions is not None and fname in self._ud_functions:
            return self._ud_functions[fname](*kwargs)
        return None

    # ---- properties ----
    def add_element_property(self, pname, ptype, pmap):
        """ptype: 'uniform' | 'region' | 'elem'."""
        if self._element_properties is None:
            self._element_properties = {}
        self._element_properties[pname] = {'type': ptype, 'idmap': pmap}

    def add_nodal_property(self, pname, ptype, pmap):
        """ptype: 'uniform' | 'region' | 'nodal'."""
        if self._nodal_properties is None:
            self._nodal_properties = {}
        self._nodal_properties[pname] = {'type': ptype, 'idmap': pmap}

    def remove_element_property(self, pname):
        if self._element_properties is not None:
            self._element_properties.pop(pname, None)
            if not self._element_properties:
                self._element_properties = None

    def remove_nodal_property(self, pname):
        if self._nodal_properties is not None:
            self._nodal_properties.pop(pname, None)
            if not self._nodal_properties:
                self._nodal_properties = None

    def remove_all_element_properties(self):
        self._element_properties = None

    def remove_all_nodal_properties(self):
        self._nodal_properties = None

    def ElementProperty(self, pname, elemtype, elemID, regionID):
        p = self._element_properties[pname]
        if p['type'] == 'uniform':
            return p['idmap']
        if p['type'] == 'region':
            return p['idmap'][regionID]
        if p['type'] == 'elem':
            return p['idmap'][elemtype][elemID]
        raise Exception('{}: unknown type'.format(p['type']))

    def NodalProperty(self, pname, pointID, regionID):
        p = self._nodal_properties[pname]
        if p['type'] == 'uniform':
            return p['idmap']
        if p['type'] == 'region':
            return p['idmap'][regionID]
        if p['type'] == 'nodal':
            return p['idmap'][pointID]
        raise Exception('{}: unknown type'.format(p['type']))

    def element_property_names(self):
        return list(self._element_properties.keys()) if self._element_properties is not None else None

    def nodal_property_names(self):
        return list(self._nodal_properties.keys()) if self._nodal_properties is not None else None

    def element_property_type(self, pname):
        return self._element_properties[pname]['type'] if self._element_properties and pname in self._element_properties else None

    def nodal_property_type(self, pname):
        return self._nodal_properties[pname]['type'] if self._nodal_properties and pname in self._nodal_properties else None
